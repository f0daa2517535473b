"""Golden fixture of the UL-SCH data path with multiplexed UCI from the reference's OWN compiled code
(oracle/_ref: srslte_ulsch_encode -> noisy LLRs -> srslte_ulsch_decode, lib/src/phy/phch/sch.c:1013-1232).

Run in the dev container:  python tests/golden/make_golden_ulsch_uci.py  ->  tests/golden/ulsch_uci_vectors.npz
Per case: the harness configuration u[12] (oracle/ref_harness.c:fill_ul_cfg), the descrambled LLRs q, the scrambling
bits c, the reference's de-multiplexed g array, srslte_ulsch_decode's return value, the decoded bytes and the average
number of half iterations.  Cases cover QPSK / 16QAM / 64QAM, 12 and 10 PUSCH symbols, ACK of 1 / 2 / 4 bits, RI of
1 / 2 bits (the 1-bit RI decoder writes into q, visible through the g[0] quirk for QPSK), short and long CQI."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_libs as ol  # noqa: E402

R, P = ol.ref(), ol.port()
assert R is not None, "build oracle/_ref first (make -C oracle ref)"
rng = np.random.default_rng(20261019)
t = R.refh_tb_new()
out = {}
cases = [  # tbs, qm, L_prb, nof_symb, nof_ack, ri_len, cqi_mode
    (1000, 2, 6, 12, 1, 1, 0), (1000, 2, 6, 12, 2, 2, 1), (936, 2, 6, 10, 0, 1, 2),
    (2792, 4, 6, 12, 1, 0, 0), (2792, 4, 6, 12, 0, 1, 0), (2792, 4, 6, 12, 2, 2, 2), (1544, 4, 6, 10, 4, 1, 1),
    (4008, 6, 6, 12, 1, 1, 1), (4008, 6, 6, 12, 2, 0, 2), (2600, 6, 5, 10, 2, 2, 0),
    (5736, 4, 12, 12, 2, 1, 2), (1000, 2, 6, 12, 0, 0, 1),
]
for ci, (tbs, qm, l_prb, nsymb, nof_ack, ri_len, cqi) in enumerate(cases):
    nb_q = qm * l_prb * 12 * nsymb
    u = ol.ul_cfg(tbs, qm, 0, nb_q, l_prb, nsymb, nof_ack, ri_len, cqi, i_ack=int(rng.integers(0, 12)),
                  i_ri=int(rng.integers(0, 13)), i_cqi=int(rng.integers(2, 16)))
    data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
    qb = np.zeros(nb_q, np.uint8)
    assert R.refh_ulsch_encode(t, u, data, np.array([1, 0, 1, 1], np.uint8), 1, qb) >= 0
    c = np.zeros(nb_q, np.uint8)
    P.port_gold_sequence(int(rng.integers(1, 2 ** 31 - 1)), nb_q, c)
    sigma = (0.25, 0.45, 0.6)[ci % 3]
    llr = np.clip(((2.0 * qb - 1) + sigma * rng.standard_normal(nb_q)) * 300, -32000, 32000).astype(np.int16)
    g = np.zeros(nb_q, np.int16)
    dec = np.zeros(tbs // 8 + 8, np.uint8)
    avg = C.c_float()
    uo = np.zeros(4, np.uint8)
    R.refh_tb_rx_reset(t, tbs)
    rc = R.refh_ulsch_decode(t, u, llr, c, g, dec, 10, C.byref(avg), uo)
    out[f"c{ci}_u"] = u
    out[f"c{ci}_q"] = llr
    out[f"c{ci}_c"] = np.packbits(c)
    out[f"c{ci}_g"] = g
    out[f"c{ci}_ret"] = np.array([rc], np.int64)
    out[f"c{ci}_dec"] = dec[: tbs // 8 + 3]
    out[f"c{ci}_avg"] = np.array([avg.value], np.float32)
    out[f"c{ci}_data"] = data
    print(ci, list(u), "ret", rc, "avg", avg.value, "ok", np.array_equal(dec[: tbs // 8], data))
np.savez_compressed(os.path.join(HERE, "ulsch_uci_vectors.npz"), **out)
print("wrote", len(out), "arrays")
