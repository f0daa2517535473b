"""The oracle port side by side with the reference's own compiled code (oracle/_ref/libsrslte_ref.so).

Skipped when oracle/_ref has not been built (it is built in the dev container, where /root/reference
exists, by oracle/Makefile, and travels to the GPU box as a prebuilt file).  CPU only.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_libs as ol

pytestmark = pytest.mark.skipif(ol.ref() is None, reason="oracle/_ref not built")


@pytest.mark.parametrize("K", [40, 104, 400, 408, 512, 800, 816, 1024, 2048, 5824, 6144])
def test_decoder_matches_reference(K, vec):
    for si, (sigma, scale) in enumerate([(1.457, 100), (1.092, 100), (0.9, 700), (0.6, 4000)]):
        bits, llr = vec.make_blocks(3, K, sigma, scale, seed=K + si)
        for nit in (1, 2, 3, 4, 7, 10):
            assert np.array_equal(ol.port_run_all(llr, K, nit), ol.ref_run_all(llr, K, nit)), (K, sigma, scale, nit)
        sb = vec.sb_layout_from_natural(llr, K)
        assert np.array_equal(ol.port_run_all(sb, K, 5, natural=False), ol.ref_run_all(sb, K, 5, natural=False))


def test_soft_outputs_match_reference(vec):
    for K in (408, 1024, 6144):
        bits, llr = vec.make_blocks(1, K, 1.092, 100, seed=K)
        pb, ps, _ = ol.port_trace(llr[0], K, 6)
        rb, rs = ol.ref_trace(llr[0], K, 6)
        assert np.array_equal(pb, rb) and np.array_equal(ps, rs)


def test_all_188_sizes_one_block(vec):
    for K in ol.ALL_K:
        bits, llr = vec.make_blocks(1, K, 1.092, 100, seed=K)
        assert np.array_equal(ol.port_run_all(llr, K, 3), ol.ref_run_all(llr, K, 3)), K


def test_rm_rx_matches_reference():
    P, R = ol.port(), ol.ref()
    rng = np.random.default_rng(11)
    for K in ol.ALL_K[::9] + [6144]:
        idx = ol.ALL_K.index(K)
        for rv in range(4):
            for E in (K, 3 * K + 12, 7 * K + 5):
                e = rng.integers(-3000, 3000, E).astype(np.int16)
                for sb in (True, False):
                    a = rng.integers(-30000, 30000, 18600).astype(np.int16)
                    b = a.copy()
                    assert P.port_rm_turbo_rx(e, E, a, K, rv, int(sb)) == 0
                    assert R.srslte_rm_turbo_rx_lut_(e.copy(), b, E, idx, rv, sb) == 0
                    assert np.array_equal(a, b), (K, rv, E, sb)


def test_cbsegm_matches_reference():
    P, R = ol.port(), ol.ref()
    for tbs in list(range(16, 6200, 8)) + list(range(6200, 100000, 136)) + [75376]:
        a, b = ol.PortCbsegm(), ol.PortCbsegm()
        assert P.port_cbsegm(C.byref(a), tbs) == R.srslte_cbsegm(C.byref(b), tbs)
        assert bytes(a) == bytes(b), tbs


def test_encoder_mirror_matches_reference(vec, golden):
    """the TX mirror in vectors.py equals srslte_tcod_encode / srslte_rm_turbo_tx (vector generation only)."""
    R = ol.ref()
    rng = np.random.default_rng(2)
    for K in (40, 504, 6144):
        b = rng.integers(0, 2, (2, K), dtype=np.uint8)
        c = vec.turbo_encode(b)
        for i in range(2):
            o = np.zeros(3 * K + 12, np.uint8)
            assert R.refh_tcod_encode(b[i].copy(), o, K) == 0
            assert np.array_equal(o, c[i])
        for rv in range(4):
            for E in (100, 3 * K + 12, 4 * K + 77):
                o = np.zeros(E, np.uint8)
                assert R.refh_rm_turbo_tx(c[0].copy(), K, o, E, rv) == 0
                assert np.array_equal(o, vec.rate_match(c[:1], E, rv)[0]), (K, rv, E)
    # the reference's stored K=504 code word vs its own encoder: one differing tail bit
    k = golden["kat"]
    o = np.zeros(1524, np.uint8)
    R.refh_tcod_encode(k["known_data"].copy(), o, 504)
    assert np.nonzero(o != k["known_data_encoded"])[0].tolist() == [1512]
