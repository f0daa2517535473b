"""Front end on the GPU (soft demodulation + descrambling, alone and fused into rate de-matching) against the
oracle port and the golden vectors of the compiled reference.  B200 only."""
import os

import numpy as np
import pytest

import oracle_libs as ol

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frontend_vectors.npz")


def _run(ctx, cws, syms):
    import torch
    s = torch.from_numpy(np.ascontiguousarray(np.concatenate(syms)).view(np.float32)).cuda()
    total = sum(c["qm"] * c["nof_symbols"] for c in cws)
    e = torch.zeros(total, dtype=torch.int16, device="cuda")
    ctx.demod_descramble_dev(cws, s.data_ptr(), e.data_ptr())
    ctx.synchronize()
    return e.cpu().numpy()


def test_golden_vectors_of_the_reference(ctx):
    g = np.load(GOLD)
    cws, syms, want, so, lo = [], [], [], 0, 0
    n = 0
    while f"c{n}_par" in g:
        qm, nsym, c_init, nb = (int(v) for v in g[f"c{n}_par"])
        cws.append(dict(qm=qm, nof_symbols=nsym, c_init=c_init, nof_bits=nb, sym_offset=so, llr_offset=lo))
        syms.append(g[f"c{n}_sym"]); want.append(g[f"c{n}_llr"])
        so += nsym; lo += qm * nsym
        n += 1
    got = _run(ctx, cws, syms)
    assert np.array_equal(got, np.concatenate(want))


def test_many_codewords_vs_oracle(ctx):
    """A batch shaped like BASELINE config 5 (many UEs per subframe, mixed modulations and sizes) plus a full
    20 MHz 64QAM codeword (config 2: 15000 symbols, 90000 bits)."""
    rng = np.random.default_rng(11)
    cws, syms, want, so, lo = [], [], [], 0, 0
    shapes = [(6, 15000)] + [(int(rng.choice([2, 4, 6, 8])), int(rng.integers(1, 3000))) for _ in range(60)]
    for qm, nsym in shapes:
        amp = float(rng.choice([0.3, 1.0, 2.0]))
        sym = ((rng.standard_normal(nsym) + 1j * rng.standard_normal(nsym)) * amp).astype(np.complex64)
        c_init = int(rng.integers(1, 2 ** 31 - 1))
        nb = qm * nsym - int(rng.integers(0, min(qm * nsym, 13)))
        cws.append(dict(qm=qm, nof_symbols=nsym, c_init=c_init, nof_bits=nb, sym_offset=so, llr_offset=lo))
        syms.append(sym); want.append(ol.port_demod_descramble(qm, sym, c_init, nb))
        so += nsym; lo += qm * nsym
    got = _run(ctx, cws, syms)
    assert np.array_equal(got, np.concatenate(want))


def test_fused_with_rate_dematching_vs_oracle(ctx):
    """symbols -> LLR -> descramble -> srslte_rm_turbo_rx_lut in ONE kernel (no e array), with HARQ combining of two
    transmissions, against the oracle: port_demod_descramble, then the port's receive index table applied as
    work[table[i mod N]] += e[i] (wrapping int16)."""
    import torch
    P = ol.port()
    rng = np.random.default_rng(12)
    # three codewords; each carries a few code blocks the way sch.c:324-334 cuts them (E LLRs per block)
    plan = [(6, [5824, 5824, 5824], 6918), (4, [1024, 1056], 25000), (2, [40, 512, 6144], 1200)]
    wl = 18624
    want = np.zeros((8, wl), np.int64)
    work = torch.zeros((8, wl), dtype=torch.int16, device="cuda")
    for rv in (0, 2):
        cws, syms, blocks, so, bi = [], [], [], 0, 0
        for ci, (qm, Ks, E) in enumerate(plan):
            E = E // qm * qm
            nsym = len(Ks) * E // qm
            sym = ((rng.standard_normal(nsym) + 1j * rng.standard_normal(nsym)) * 0.8).astype(np.complex64)
            c_init = int(rng.integers(1, 2 ** 31 - 1))
            cws.append(dict(qm=qm, nof_symbols=nsym, c_init=c_init, sym_offset=so))
            syms.append(sym)
            so += nsym
            e = ol.port_demod_descramble(qm, sym, c_init).astype(np.int64)
            for j, K in enumerate(Ks):
                blocks.append((K, rv, ci, j * E, E, bi * wl))
                tab = np.zeros(3 * K + 12, np.uint16)
                assert P.port_rm_rx_table(K, rv, 1, tab) == 0
                np.add.at(want[bi], tab[np.arange(E) % (3 * K + 12)].astype(np.int64), e[j * E:(j + 1) * E])
                bi += 1
        s = torch.from_numpy(np.ascontiguousarray(np.concatenate(syms)).view(np.float32)).cuda()
        ctx.demod_rm_rx_batch_dev(cws, blocks, s.data_ptr(), work.data_ptr())
        ctx.synchronize()
        assert np.array_equal(work.cpu().numpy(), want.astype(np.int16)), rv   # astype wraps like the int16 "+="


def test_argument_errors(ctx):
    import torch
    s = torch.zeros(64, dtype=torch.float32, device="cuda")
    e = torch.zeros(64, dtype=torch.int16, device="cuda")
    with pytest.raises(Exception):
        ctx.demod_descramble_dev([dict(qm=3, nof_symbols=4, c_init=1)], s.data_ptr(), e.data_ptr())
    with pytest.raises(Exception):
        ctx.demod_descramble_dev([dict(qm=2, nof_symbols=4, c_init=1, nof_bits=9)], s.data_ptr(), e.data_ptr())
