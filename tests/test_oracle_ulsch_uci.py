"""UL-SCH data path with multiplexed UCI (srslte_ulsch_decode, lib/src/phy/phch/sch.c:920-1064):
the oracle port against the committed golden vectors of the compiled reference and, when oracle/_ref is present,
against the reference itself.  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_libs as ol

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ulsch_uci_vectors.npz")


def _cmp_g(g_port, g_ref, q_cqi, qm):
    # decode_cqi_short accumulates the repetitions of the CQI code word into its first 32 LLRs in place (uci.c:317-322),
    # so the CQI part of the reference's g array is only comparable when it has no repetition
    eo = q_cqi * qm if q_cqi * qm > 32 else 0
    return np.array_equal(g_port[eo:], g_ref[eo:g_port.size])


def test_port_matches_golden_ulsch_uci():
    g = np.load(GOLD)
    n = n_ok = 0
    seen_g0 = 0
    while f"c{n}_u" in g:
        u = g[f"c{n}_u"]
        c = np.unpackbits(g[f"c{n}_c"])[: int(u[3])]
        rc, dec, avg, gp, (qa, qr, qc) = ol.port_ulsch_decode(u, g[f"c{n}_q"], c, 10)
        assert _cmp_g(gp, g[f"c{n}_g"], qc, int(u[1])), n
        assert rc == int(g[f"c{n}_ret"][0]), n
        assert np.array_equal(dec, g[f"c{n}_dec"]), n
        assert abs(avg - float(g[f"c{n}_avg"][0])) < 1e-6, n
        if rc == 0:
            n_ok += 1
            assert np.array_equal(dec[: int(u[0]) // 8], g[f"c{n}_data"])
        if qr and qc == 0:   # the RI sample the reference leaves in g[0]
            seen_g0 += int(gp[0] != ol.port_ulsch_deinterleave(g[f"c{n}_q"], int(u[1]), int(u[5]))[0])
        n += 1
    assert n == 12 and n_ok >= 8 and seen_g0 >= 2


def test_q_prime_helpers_of_the_library_match_the_port():
    import __graft_entry__ as ge
    L = ge.load_package().lib()
    P = ol.port()
    rng = np.random.default_rng(9)
    for _ in range(3000):
        O = int(rng.integers(1, 23)); K = int(rng.integers(40, 80000)); lp = int(rng.integers(1, 101)); ns = int(rng.choice([9, 10, 11, 12]))
        b = float(rng.choice(ol.BETA_HARQ[:15] + ol.BETA_RI[:13] + ol.BETA_CQI[2:]))
        assert L.srslte_b200_uci_q_prime_ri_ack(O, K, lp, ns, b) == P.port_uci_q_prime_ri_ack(O, K, lp, ns, b)
        qr = int(rng.integers(0, 5))
        assert L.srslte_b200_uci_q_prime_cqi(O, K, lp, ns, b, qr) == P.port_uci_q_prime_cqi(O, K, lp, ns, b, qr)


@pytest.mark.skipif(ol.ref() is None, reason="oracle/_ref not built (needs /root/reference)")
def test_port_matches_compiled_reference_ulsch_uci():
    R, P = ol.ref(), ol.port()
    t = R.refh_tb_new()
    rng = np.random.default_rng(5)
    n = n_ok = 0
    for (tbs, qm, l_prb, nsymb) in [(2792, 4, 6, 12), (5736, 4, 12, 12), (1000, 2, 6, 12), (14112, 6, 20, 12),
                                    (1544, 4, 6, 10), (9912, 4, 25, 12), (1000, 2, 6, 9), (4008, 6, 6, 11)]:
        for nof_ack in (0, 1, 2, 4):
            for ri_len in (0, 1, 2):
                for cqi in (0, 1, 2):
                    nb_q = qm * l_prb * 12 * nsymb
                    u = ol.ul_cfg(tbs, qm, 0, nb_q, l_prb, nsymb, nof_ack, ri_len, cqi, i_ack=int(rng.integers(0, 15)),
                                  i_ri=int(rng.integers(0, 13)), i_cqi=int(rng.integers(2, 16)))
                    data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
                    qb = np.zeros(nb_q, np.uint8)
                    assert R.refh_ulsch_encode(t, u, data, np.array([1, 0, 1, 1], np.uint8), 1, qb) >= 0
                    c = np.zeros(nb_q, np.uint8)
                    P.port_gold_sequence(int(rng.integers(1, 2 ** 31 - 1)), nb_q, c)
                    llr = np.clip(((2.0 * qb - 1) + 0.4 * rng.standard_normal(nb_q)) * 300, -32000, 32000).astype(np.int16)
                    g = np.zeros(nb_q, np.int16)
                    out = np.zeros(tbs // 8 + 8, np.uint8)
                    avg = C.c_float()
                    R.refh_tb_rx_reset(t, tbs)
                    rc = R.refh_ulsch_decode(t, u, llr, c, g, out, 6, C.byref(avg), np.zeros(4, np.uint8))
                    prc, pdec, pavg, gp, (qa, qr, qc) = ol.port_ulsch_decode(u, llr, c, 6)
                    assert _cmp_g(gp, g, qc, qm), list(u)
                    assert prc == rc and np.array_equal(pdec, out[: tbs // 8 + 3]) and abs(pavg - avg.value) < 1e-6, list(u)
                    n += 1
                    n_ok += int(rc == 0)
    R.refh_tb_free(t)
    assert n == 288 and n_ok > 200
