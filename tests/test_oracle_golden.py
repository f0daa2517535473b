"""The oracle (oracle/tdec_port.c) against the committed golden vectors.

The vectors were produced by the reference's own compiled code (tests/golden/make_golden.py) and by the
reference tests' known-answer data (crc_test.h:37-38, turbodecoder_test.h:74-162).  CPU only.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_libs as ol


def _cases(g, prefix):
    return sorted({k.split("_")[0] for k in g if k.startswith(prefix)}, key=lambda s: int(s[1:]))


def test_crc_known_answers(golden):
    k = golden["kat"]
    nb = int(k["crc_nbits"][0])
    bits = np.unpackbits(k["crc_bits"])[:nb].copy()
    P = ol.port()
    # crc_test -n 5001 -l 24 -p 0x1864CFB / 0x1800063 -s 1  (lib/src/phy/fec/test/crc_test.h:37-38)
    assert P.port_crc_bits(ol.CRC24A, bits, nb) == 0x1C5C97 == int(k["crc24a"][0])
    assert P.port_crc_bits(ol.CRC24B, bits, nb) == 0x36D1F0 == int(k["crc24b"][0])


def test_crc_byte_and_bit_variants_agree():
    P = ol.port()
    rng = np.random.default_rng(3)
    for n in (8, 24, 40, 6144):
        bits = rng.integers(0, 2, n, dtype=np.uint8)
        for poly in (ol.CRC24A, ol.CRC24B):
            assert P.port_crc_bits(poly, bits, n) == P.port_crc_bytes(poly, np.packbits(bits), n)
    # attaching the CRC makes the check word zero (sch.c:365-378)
    payload = rng.integers(0, 2, 6120, dtype=np.uint8)
    crc = P.port_crc_bits(ol.CRC24B, payload, 6120)
    full = np.concatenate([payload, [(crc >> (23 - i)) & 1 for i in range(24)]]).astype(np.uint8)
    assert P.port_crc_bytes(ol.CRC24B, np.packbits(full), 6144) == 0


def test_known_codeword_k504(golden, vec):
    """turbodecoder_test.h known_data/known_data_encoded: encoder KAT and noiseless decode of it."""
    k = golden["kat"]
    data, enc = k["known_data"], k["known_data_encoded"]
    # The reference's stored code word differs from what its OWN srslte_tcod_encode produces in exactly one
    # position, the first tail bit (index 3K = 1512); oracle/_ref reproduces that (test_oracle_vs_ref.py).
    mism = np.nonzero(vec.turbo_encode(data[None, :])[0] != enc)[0]
    assert mism.tolist() == [1512]
    llr = ((enc.astype(np.int16) * 2 - 1) * 100)[None, :]
    for nit in (1, 2, 5):
        out = ol.port_run_all(llr, 504, nit)
        assert np.array_equal(np.unpackbits(out[0]), data)


def test_decoder_golden_vectors(golden):
    """decoded bytes after every half iteration 1..10 equal the compiled reference's, all three regimes."""
    g = golden["tdec_vectors"]
    for c in _cases(g, "c"):
        K = int(g[f"{c}_K"][0])
        llr, dec = g[f"{c}_llr"], g[f"{c}_dec"]
        for i in range(llr.shape[0]):
            by, _, _ = ol.port_trace(llr[i], K, 10)
            assert np.array_equal(by, dec[i]), (c, K, i)
            # run_all(n) ends on the same bytes as n successive iterations
            for nit in (1, 4, 7):
                assert np.array_equal(ol.port_run_all(llr[i:i + 1], K, nit)[0], dec[i, nit - 1]), (c, K, nit)


def test_saturating_case_present(golden):
    """at LLR scale 700 the reference's saturating adds really clamp: the fixture must exercise that."""
    g = golden["tdec_vectors"]
    clamps = 0
    for c in _cases(g, "c"):
        K, scale = int(g[f"{c}_K"][0]), int(g[f"{c}_K"][1])
        if scale == 700 and K > 400:
            _, _, n = ol.port_trace(g[f"{c}_llr"][0], K, 10)
            clamps += n
    assert clamps > 0


def test_rm_rx_tables_golden(golden):
    g = golden["rm_tables"]
    P = ol.port()
    for key, want in g.items():
        K, rv, sb = (int(x) for x in key.replace("K", "").replace("rv", "").replace("sb", "").split("_"))
        tab = np.zeros(3 * K + 12, np.uint16)
        assert P.port_rm_rx_table(K, rv, sb, tab) == 0
        assert np.array_equal(tab, want), key


def test_rm_rx_accumulates_and_wraps():
    P = ol.port()
    K, rv = 512, 1
    N = 3 * K + 12
    tab = np.zeros(N, np.uint16)
    P.port_rm_rx_table(K, rv, 1, tab)
    rng = np.random.default_rng(5)
    E = 2 * N + 77
    e = rng.integers(-20000, 20000, E).astype(np.int16)
    buf = rng.integers(-20000, 20000, ol.port().port_cb_size(0) * 0 + 18600).astype(np.int16)
    want = buf.astype(np.int64).copy()
    for i in range(E):
        want[tab[i % N]] += int(e[i])
    want = ((want + 32768) % 65536 - 32768).astype(np.int16)
    assert P.port_rm_turbo_rx(e, E, buf, K, rv, 1) == 0
    assert np.array_equal(buf, want)


def test_cbsegm_examples():
    P = ol.port()
    s = ol.PortCbsegm()
    assert P.port_cbsegm(C.byref(s), 75376) == 0  # 100 PRB MCS 28
    assert (s.C, s.K1, s.C1, s.C2, s.F) == (13, 5824, 13, 0, 0)
    assert P.port_cbsegm(C.byref(s), 6120) == 0
    assert (s.C, s.K1, s.F) == (1, 6144, 0)
    assert P.port_cbsegm(C.byref(s), 0) == 0 and s.C == 0


def test_window_rule():
    P = ol.port()
    assert [P.port_nof_subblocks(k) for k in (40, 400, 408, 800, 816, 1024, 6144)] == [0, 0, 8, 8, 16, 16, 16]
    assert len(ol.ALL_K) == 188 and all(P.port_cb_size(i) == k for i, k in enumerate(ol.ALL_K))


def test_transport_block_golden(golden):
    """decode_tb semantics incl. HARQ soft combining over rv 0 -> rv 2 and the skip-good-CB path."""
    g = golden["tb_vectors"]
    P = ol.port()
    dec = P.port_tdec_new()
    for c in _cases(g, "t"):
        tbs, qm, G, max_it = (int(x) for x in g[f"{c}_par"])
        seg = ol.PortCbsegm()
        P.port_cbsegm(C.byref(seg), tbs)
        sb = ol.PortSoftbuffer()
        assert P.port_softbuffer_init(C.byref(sb), seg.C) == 0
        for rv in (0, 2):
            out = np.zeros(tbs // 8 + 8, np.uint8)
            avg = C.c_float()
            noi = np.zeros(max(seg.C, 1), np.uint32)
            rc = P.port_decode_tb(dec, C.byref(sb), tbs, qm, rv, G, g[f"{c}_rv{rv}_llr"], out, max_it, C.byref(avg), noi)
            want_rc, want_its = (int(x) for x in g[f"{c}_rv{rv}_res"])
            assert rc == want_rc, (c, rv)
            assert round(avg.value * seg.C) == want_its == int(noi.sum()), (c, rv)
            assert np.array_equal(out[: tbs // 8 + 3], g[f"{c}_rv{rv}_out"]), (c, rv)
            assert np.array_equal(np.ctypeslib.as_array(sb.cb_crc, (seg.C,)), g[f"{c}_rv{rv}_cbcrc"]), (c, rv)
            if rc == 0:
                assert np.array_equal(out[: tbs // 8], g[f"{c}_data"])
        P.port_softbuffer_free(C.byref(sb))
    P.port_tdec_free(dec)
