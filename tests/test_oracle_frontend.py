"""Front end (soft demodulation + descrambling): the oracle port against the committed golden
vectors of the compiled reference, against the reference itself when oracle/_ref is present, and against the
36.211 definition of the scrambling sequence.  CPU only."""
import os

import numpy as np
import pytest

import oracle_libs as ol

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frontend_vectors.npz")


@pytest.fixture(scope="module")
def g():
    return np.load(GOLD)


def test_port_matches_golden_llrs(g):
    n = 0
    while f"c{n}_par" in g:
        qm, nsym, c_init, nb = (int(v) for v in g[f"c{n}_par"])
        got = ol.port_demod_descramble(qm, g[f"c{n}_sym"], c_init, nb)
        assert np.array_equal(got, g[f"c{n}_llr"]), (n, qm, nsym)
        n += 1
    assert n == 20


def test_scrambling_sequence_matches_golden_and_definition(g):
    P = ol.port()
    for si in range(3):
        seed = int(g[f"seq{si}_seed"][0])
        want = g[f"seq{si}_bits"]
        c = np.zeros(want.size, np.uint8)
        P.port_gold_sequence(seed, want.size, c)
        assert np.array_equal(c, want), si
        # 36.211 7.2 straight from the text
        x1 = np.zeros(1600 + want.size + 31, np.uint8); x2 = np.zeros_like(x1)
        x1[0] = 1
        for i in range(31):
            x2[i] = (seed >> i) & 1
        for i in range(1600 + want.size):
            x1[i + 31] = x1[i + 3] ^ x1[i]
            x2[i + 31] = x2[i + 3] ^ x2[i + 2] ^ x2[i + 1] ^ x2[i]
        assert np.array_equal((x1 ^ x2)[1600:1600 + want.size], want), si


def test_port_matches_compiled_reference():
    if ol.ref() is None:
        pytest.skip("oracle/_ref is not built (the reference tree is absent)")
    rng = np.random.default_rng(7)
    for qm in (2, 4, 6, 8):
        for n in (1, 3, 4, 7, 8, 9, 15, 16, 17, 100, 1203, 15000):
            for amp in (0.1, 1.0, 3.0, 10.0):
                sym = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * amp).astype(np.complex64)
                c_init = int(rng.integers(1, 2 ** 31 - 1))
                nb = qm * n if n % 2 else qm * n - (qm * n) // 3
                assert np.array_equal(ol.port_demod_descramble(qm, sym, c_init, nb),
                                      ol.ref_demod_descramble(qm, sym, c_init, nb)), (qm, n, amp)


def test_ulsch_deinterleaver_port_vs_reference_and_definition():
    """36.212 5.2.2.8 without UCI: g[(j*cols + i)*Qm + k] = q[(i*rows + j)*Qm + k]; the port against that formula and,
    when oracle/_ref is present, against the reference's own ulsch_deinterleave (sch.c:891-918)."""
    rng = np.random.default_rng(9)
    have_ref = ol.ref() is not None
    for qm in (2, 4, 6):
        for cols in (12, 11, 10, 9):
            for prb in (1, 3, 25, 100):
                H = prb * 12 * cols
                rows = H // cols
                q = rng.integers(-3000, 3000, H * qm).astype(np.int16)
                g = ol.port_ulsch_deinterleave(q, qm, cols)
                want = q.reshape(cols, rows, qm).transpose(1, 0, 2).reshape(-1)
                assert np.array_equal(g, want), (qm, cols, prb)
                if have_ref:
                    assert np.array_equal(g, ol.ref_ulsch_deinterleave(q, qm, cols)), (qm, cols, prb)
