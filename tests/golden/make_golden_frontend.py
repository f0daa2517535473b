"""Golden fixture of the front end (soft demodulation + descrambling) from the reference's OWN
compiled code (oracle/_ref: srslte_demod_soft_demodulate_s + srslte_sequence_LTE_pr + srslte_scrambling_s_offset).

Run in the dev container:  python tests/golden/make_golden_frontend.py  ->  tests/golden/frontend_vectors.npz
Cases: every modulation (QPSK, 16QAM, 64QAM, 256QAM) x symbol counts that exercise the SIMD bodies and the scalar
remainders of the reference (multiples of 4 / 8 and not), partial descrambling lengths, and the first 256 bits of
the scrambling sequence for three seeds (obtained by descrambling a constant +1 LLR vector)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_libs as ol  # noqa: E402

assert ol.ref() is not None, "build oracle/_ref first (make -C oracle ref)"
rng = np.random.default_rng(20261018)
out = {}
ci = 0
for qm in (2, 4, 6, 8):
    for n in (5, 8, 19, 64, 257):
        amp = (0.4, 1.0, 2.5)[ci % 3]
        sym = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * amp).astype(np.complex64)
        c_init = int(rng.integers(1, 2 ** 31 - 1))
        nb = qm * n if ci % 2 == 0 else qm * n - 7
        out[f"c{ci}_par"] = np.array([qm, n, c_init, nb], np.int64)
        out[f"c{ci}_sym"] = sym
        out[f"c{ci}_llr"] = ol.ref_demod_descramble(qm, sym, c_init, nb)
        ci += 1
# the scrambling sequence itself: QPSK of symbols whose LLRs are a constant, sign = 1 - 2 c(n)
for si, seed in enumerate((1, 0x12345678 & 0x7FFFFFFF, (0x46 << 14) + (3 << 9) + 1)):
    sym = np.full(128, -(1 + 1j) / np.sqrt(2) / 100 * 50, np.complex64)   # LLR = +50 before descrambling
    llr = ol.ref_demod_descramble(2, sym, seed)
    assert set(np.abs(llr)) == {50}
    out[f"seq{si}_seed"] = np.array([seed], np.int64)
    out[f"seq{si}_bits"] = (llr < 0).astype(np.uint8)
np.savez_compressed(os.path.join(HERE, "frontend_vectors.npz"), **out)
print("wrote", len(out), "arrays")
