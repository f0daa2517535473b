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


def test_aligned_codewords_take_the_vector_path(ctx):
    """Codewords whose symbols and LLRs start on 128-bit boundaries are demodulated G = 8 (24 for 64QAM) LLRs per thread;
    every modulation with symbol counts around the reference's SIMD-body / scalar-remainder boundaries, descrambled
    lengths that end inside a group or leave whole groups undescrambled, and one deliberately misaligned neighbour."""
    import torch
    rng = np.random.default_rng(23)
    cws, want, so, lo = [], [], 0, 0
    sizes = [1, 2, 3, 4, 5, 7, 8, 9, 12, 13, 15, 16, 17, 31, 33, 100, 255, 1001, 2998, 3000]
    chunks = []
    for qm in (2, 4, 6, 8):
        for nsym in sizes:
            for cut in (0, 1, 9, 30):
                n = qm * nsym
                nb = max(1, n - cut)
                amp = float(rng.choice([0.3, 1.0, 5.0]))
                sym = ((rng.standard_normal(nsym) + 1j * rng.standard_normal(nsym)) * amp).astype(np.complex64)
                c_init = int(rng.integers(1, 2 ** 31 - 1))
                mis = 1 if (nsym == 13 and cut == 1) else 0     # an odd symbol offset: the per-LLR path
                so += mis
                cws.append(dict(qm=qm, nof_symbols=nsym, c_init=c_init, nof_bits=nb, sym_offset=so, llr_offset=lo))
                chunks.append((so, sym)); want.append((lo, ol.port_demod_descramble(qm, sym, c_init, nb)))
                so = (so + nsym + 1) & ~1
                lo = (lo + n + 7) & ~7
    sbuf = np.zeros(so + 2, np.complex64)
    for o, sym in chunks:
        sbuf[o:o + len(sym)] = sym
    st = torch.from_numpy(sbuf.view(np.float32)).cuda()
    e = torch.full((lo + 8,), 12345, dtype=torch.int16, device="cuda")
    ctx.demod_descramble_dev(cws, st.data_ptr(), e.data_ptr())
    ctx.synchronize()
    got = e.cpu().numpy()
    for o, w in want:
        assert np.array_equal(got[o:o + len(w)], w)
    # nothing outside the codewords was written
    mask = np.ones(len(got), bool)
    for o, w in want:
        mask[o:o + len(w)] = False
    assert np.all(got[mask] == 12345)


def test_pusch_codewords_with_ulsch_deinterleaver_vs_oracle(ctx):
    """PUSCH order of operations (pusch.c:482-500, sch.c:1028-1036): demodulate, descramble, then the UL-SCH channel
    de-interleaver; the kernel applies the permutation as an index map."""
    rng = np.random.default_rng(14)
    cws, syms, want, so, lo = [], [], [], 0, 0
    for qm in (2, 4, 6):
        for cols in (12, 11, 10):
            for prb in (1, 6, 50):
                nsym = prb * 12 * cols
                sym = ((rng.standard_normal(nsym) + 1j * rng.standard_normal(nsym)) * 0.9).astype(np.complex64)
                c_init = int(rng.integers(1, 2 ** 31 - 1))
                cws.append(dict(qm=qm, nof_symbols=nsym, c_init=c_init, sym_offset=so, llr_offset=lo, ul_nof_symb=cols))
                syms.append(sym)
                want.append(ol.port_ulsch_deinterleave(ol.port_demod_descramble(qm, sym, c_init), qm, cols))
                so += nsym
                lo += qm * nsym
    got = _run(ctx, cws, syms)
    assert np.array_equal(got, np.concatenate(want))


@pytest.mark.parametrize("align", [False, True])
def test_fused_with_rate_dematching_vs_oracle(ctx, align):
    """symbols -> LLR -> descramble -> srslte_rm_turbo_rx_lut in ONE kernel (no e array), with HARQ combining of two
    transmissions, against the oracle: port_demod_descramble, then the port's receive index table applied as
    work[table[i mod N]] += e[i] (wrapping int16).  align: every codeword starts on a 128-bit boundary (the kernel's
    vector path for blocks without wrap-around); packed: odd symbol offsets (one LLR at a time)."""
    import torch
    P = ol.port()
    rng = np.random.default_rng(12)
    # three codewords; each carries a few code blocks the way sch.c:324-334 cuts them (E LLRs per block)
    plan = [(6, [5824, 5824, 5824], 6918), (4, [1024, 1056], 25000), (2, [40, 512, 6144], 1200),
            (4, [2048, 6144, 3072], 5004), (8, [4096, 1504], 4568), (6, [6144], 18438)]
    wl = 18624
    want = np.zeros((14, wl), np.int64)
    work = torch.zeros((14, wl), dtype=torch.int16, device="cuda")
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
            if align and so % 2:
                syms.append(np.zeros(1, np.complex64))
                so += 1
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


def _modulate(bits, qm):
    """36.211 7.1 mapping (QPSK, 16QAM, 64QAM): bits [n * qm] -> complex64 [n]."""
    b = 1.0 - 2.0 * bits.reshape(-1, qm).astype(np.float64)
    if qm == 2:
        i, q, nrm = b[:, 0], b[:, 1], np.sqrt(2)
    elif qm == 4:
        i, q, nrm = b[:, 0] * (2 - b[:, 2]), b[:, 1] * (2 - b[:, 3]), np.sqrt(10)
    else:
        i, q, nrm = b[:, 0] * (4 - b[:, 2] * (2 - b[:, 4])), b[:, 1] * (4 - b[:, 3] * (2 - b[:, 5])), np.sqrt(42)
    return ((i + 1j * q) / nrm).astype(np.complex64)


def test_transport_blocks_from_symbols_vs_oracle(ctx, vec):
    """Whole chain from equalised symbols to transport-block bytes in one call (pdsch.c:760-796): TBs are built
    with the numpy TX mirror (CRC, segmentation, turbo encoder, rate matching, scrambling, modulation, AWGN) and
    decoded by the oracle (port_demod_descramble -> port_decode_tb) and by srslte_b200_decode_tb_sym_batch."""
    import ctypes as C
    P = ol.port()
    rng = np.random.default_rng(21)
    # (tbs, qm, G, N_pusch_symbs or 0 for PDSCH)
    cases = [(2216, 2, 4800, 0), (6200, 4, 9600, 0), (14112, 4, 28800, 0), (36696, 6, 60000, 0), (75376, 6, 90000, 0),
             (1000, 2, 2400, 0), (2216, 4, 4800, 12), (14112, 4, 28800, 12), (36696, 6, 59904, 12), (6200, 4, 9504, 11)]
    descs, want = [], []
    dec = P.port_tdec_new()
    for i, (tbs, qm, G, ul) in enumerate(cases * 2):
        seg = ol.PortCbsegm()
        assert P.port_cbsegm(C.byref(seg), tbs) == 0 and seg.F == 0
        payload = rng.integers(0, 2, tbs, dtype=np.uint8)
        tb = vec.attach_crc(vec.CRC24A, payload[None, :])[0]
        e_parts, pos = [], 0
        Gp, gamma = G // qm, (G // qm) % seg.C
        for cb in range(seg.C):
            K = seg.K1 if cb < seg.C1 else seg.K2
            rlen = K if seg.C == 1 else K - 24
            blk = tb[pos:pos + rlen]
            pos += rlen
            if seg.C > 1:
                blk = vec.attach_crc(vec.CRC24B, blk[None, :])[0]
            E = qm * (Gp // seg.C) if cb <= seg.C - gamma - 1 else qm * ((Gp + seg.C - 1) // seg.C)
            e_parts.append(vec.rate_match(vec.turbo_encode(blk[None, :]), E, 0)[0])
        e = np.concatenate(e_parts).astype(np.uint8)
        assert e.size == G
        c_init = int(rng.integers(1, 2 ** 31 - 1))
        c = np.zeros(G, np.uint8)
        P.port_gold_sequence(c_init, G, c)
        sigma = (0.0, 0.08, 0.25)[i % 3]
        if ul:   # UL-SCH channel interleaver (36.212 5.2.2.8): q[(i*rows + j)*Qm + k] = g[(j*cols + i)*Qm + k]
            rows = G // qm // ul
            tx = e.reshape(rows, ul, qm).transpose(1, 0, 2).reshape(-1)
        else:
            tx = e
        sym = _modulate(tx ^ c, qm)
        sym = (sym + sigma * (rng.standard_normal(sym.size) + 1j * rng.standard_normal(sym.size))).astype(np.complex64)
        llr = ol.port_demod_descramble(qm, sym, c_init, G)
        if ul:
            llr = ol.port_ulsch_deinterleave(llr, qm, ul)
        sb = ol.PortSoftbuffer()
        P.port_softbuffer_init(C.byref(sb), seg.C)
        out = np.zeros(tbs // 8 + 8, np.uint8)
        avg = C.c_float()
        noi = np.zeros(seg.C, np.uint32)
        rc = P.port_decode_tb(dec, C.byref(sb), tbs, qm, 0, G, llr, out, 8, C.byref(avg), noi)
        want.append((rc, out[: tbs // 8 + 3].copy(), avg.value, np.packbits(payload)))
        P.port_softbuffer_free(C.byref(sb))
        descs.append(dict(tbs=tbs, qm=qm, rv=0, nof_e_bits=G, softbuffer=i, c_init=c_init, symbols=sym, ul_nof_symb=ul))
    P.port_tdec_free(dec)
    pool = ctx.harq_pool(len(descs), 13)
    got = ctx.decode_tb_sym_batch(pool, descs, 8)
    n_ok = 0
    for i, ((ret, data, avg), (rc, out, wavg, payload)) in enumerate(zip(got, want)):
        tbs = descs[i]["tbs"]
        assert ret == rc, i
        assert abs(avg - wavg) < 1e-6, (i, avg, wavg)
        assert np.array_equal(data[: tbs // 8 + 3], out), i
        if ret == 0:
            n_ok += 1
            assert np.array_equal(data[: tbs // 8], payload)
    assert n_ok >= 13
    pool.close()


def test_tx_mirror_encoder_and_rate_matching(ctx, vec):
    """Device TX mirror (turbo encoder + rate matching) against the numpy mirror, which the CPU suite pins to
    srslte_tcod_encode / srslte_rm_turbo_tx; then the round trip through the device decoder."""
    import torch
    rng = np.random.default_rng(41)
    blocks, want, bits_all, bo, eo = [], [], [], 0, 0
    for K in (40, 104, 504, 1024, 5824, 6144):
        for rv in range(4):
            for E in (96, 3 * K + 12, 4 * K + 77):
                b = rng.integers(0, 2, K, dtype=np.uint8)
                blocks.append((K, rv, E, bo, eo))
                bits_all.append(b)
                want.append(vec.rate_match(vec.turbo_encode(b[None, :]), E, rv)[0])
                bo += K
                eo += E
    bits_d = torch.from_numpy(np.concatenate(bits_all)).cuda()
    e_d = torch.zeros(eo, dtype=torch.uint8, device="cuda")
    ctx.tcod_rm_tx_batch_dev(blocks, bits_d.data_ptr(), e_d.data_ptr())
    ctx.synchronize()
    assert np.array_equal(e_d.cpu().numpy(), np.concatenate(want))
    # round trip: encode 64 blocks of K = 6144 on the device (rv 0, whole circular buffer in natural order is not what
    # the selection gives, so go through the rate de-matcher), decode, compare
    K, n = 6144, 64
    bits = rng.integers(0, 2, (n, K), dtype=np.uint8)
    N = 3 * K + 12
    bd = torch.from_numpy(bits.reshape(-1)).cuda()
    ed = torch.zeros(n * N, dtype=torch.uint8, device="cuda")
    ctx.tcod_rm_tx_batch_dev([(K, 0, N, i * K, i * N) for i in range(n)], bd.data_ptr(), ed.data_ptr())
    llr = ((ed.to(torch.int16) * 2 - 1) * 100).contiguous()
    wl = 18624
    work = torch.zeros((n, wl), dtype=torch.int16, device="cuda")
    ctx.rm_rx_batch_dev([(K, 0, i * N, N, i * wl) for i in range(n)], llr.data_ptr(), work.data_ptr())
    ctx.synchronize()
    got, _, _ = ctx.tdec_batch_host(work.cpu().numpy(), K, 2, natural=False)
    assert np.array_equal(got, np.packbits(bits, axis=1))


# ---- PUSCH with multiplexed UCI (data path of srslte_ulsch_decode, sch.c:920-1064) ----------
def _uci_positions(is_ri, n, qm, rows, cols):
    """channel positions of the n coded ACK / RI symbols (uci.c:497-545), [n, qm]."""
    sets = {(0, True): (2, 3, 8, 9), (0, False): (1, 2, 6, 7), (1, True): (1, 4, 7, 10), (1, False): (0, 3, 5, 8)}
    s = sets[(int(is_ri), cols > 10)]
    idx = np.arange(n)
    row = rows - 1 - idx // 4
    col = np.array([s[(3 * i) % 4] for i in idx], dtype=np.int64)
    return (row[:, None] * qm + rows * col[:, None] * qm + np.arange(qm)[None, :]).astype(np.int64)


def test_pusch_codewords_with_multiplexed_uci_vs_oracle(ctx):
    """De-multiplexing as an index map: ACK erasure, RI skipping with the reference's g[0] behaviour, 12 / 11 / 10 / 9
    PUSCH symbols, RI rows that reach up the matrix, against the sequential port (pinned to the compiled reference by
    tests/test_oracle_ulsch_uci.py)."""
    rng = np.random.default_rng(77)
    P = ol.port()
    cws, syms, want, so, lo = [], [], [], 0, 0
    shapes = []
    for qm in (2, 4, 6):
        for cols in (12, 11, 10, 9):
            for l_prb in (1, 3, 8):
                rows = l_prb * 12
                for q_ack, q_ri, ri_len in ((0, 1, 1), (3, 0, 0), (5, 2, 1), (17, 9, 2), (4 * rows, 7, 1), (2, 4 * rows, 2),
                                            (0, 4 * rows - 3, 1), (11, 4, 1)):
                    shapes.append((qm, cols, rows, q_ack, q_ri, ri_len))
    for qm, cols, rows, q_ack, q_ri, ri_len in shapes:
        nsym = rows * cols
        sym = ((rng.standard_normal(nsym) + 1j * rng.standard_normal(nsym)) * 0.9).astype(np.complex64)
        c_init = int(rng.integers(1, 2 ** 31 - 1))
        cws.append(dict(qm=qm, nof_symbols=nsym, c_init=c_init, sym_offset=so, llr_offset=lo, ul_nof_symb=cols,
                        uci=dict(q_prime_ack=q_ack, q_prime_ri=q_ri, q_prime_cqi=0, ri_len=ri_len)))
        q = ol.port_demod_descramble(qm, sym, c_init)
        c = np.zeros(qm * nsym, np.uint8)
        P.port_gold_sequence(c_init, qm * nsym, c)
        g = ol.port_ulsch_demux(q, c, qm, cols, q_ack, q_ri, ri_len)
        syms.append(sym)
        want.append(np.concatenate([g, np.zeros(q_ri * qm, np.int16)]))   # the undefined tail reads 0 here
        so += nsym; lo += qm * nsym
    got = _run(ctx, cws, syms)
    wantc = np.concatenate(want)
    if not np.array_equal(got, wantc):
        lo = 0
        for (qm, cols, rows, q_ack, q_ri, ri_len), w in zip(shapes, want):
            bad = np.nonzero(got[lo:lo + w.size] != w)[0]
            assert bad.size == 0, (qm, cols, rows, q_ack, q_ri, ri_len, bad[:8], got[lo + bad[:8]], w[bad[:8]])
            lo += w.size
    # more ACK / RI symbols than 4 per matrix row: the reference fails (uci.c:504, 529), so does the library
    import torch
    s = torch.zeros(2 * 144, dtype=torch.float32, device="cuda")
    e = torch.zeros(2 * 144, dtype=torch.int16, device="cuda")
    with pytest.raises(RuntimeError):
        ctx.demod_descramble_dev([dict(qm=2, nof_symbols=144, c_init=1, ul_nof_symb=12, uci=dict(q_prime_ri=49))],
                                 s.data_ptr(), e.data_ptr())


def test_pusch_transport_blocks_with_uci_from_symbols_vs_oracle(ctx, vec):
    """srslte_ulsch_decode's data path from equalised symbols to bytes in one call: the TX side (numpy mirror) puts CQI
    bits in front of the rate-matched data, interleaves around the RI cells and overwrites the ACK cells; the oracle is
    port_demod_descramble -> port_ulsch_demux -> port_decode_tb; also the committed golden cases of the compiled
    reference, fed as QPSK-like symbols is not possible (they are LLRs), so they pin the port in the CPU suite."""
    import ctypes as C
    P = ol.port()
    rng = np.random.default_rng(31)
    # tbs, qm, L_prb, N_pusch_symbs, nof_ack, ri_len, cqi_mode
    cases = [(2216, 4, 10, 12, 1, 1, 1), (1000, 2, 6, 12, 2, 1, 0), (14112, 4, 50, 12, 2, 2, 2), (36696, 6, 100, 12, 1, 1, 2),
             (6200, 4, 20, 11, 4, 1, 1), (2600, 6, 6, 10, 0, 2, 1), (4008, 6, 8, 12, 2, 0, 0), (9912, 4, 25, 9, 1, 1, 0),
             (75376, 6, 100, 12, 2, 1, 1)]
    descs, want = [], []
    for i, (tbs, qm, l_prb, cols, nof_ack, ri_len, cqi) in enumerate(cases * 2):
        rows = l_prb * 12
        nb_q = qm * rows * cols
        u = ol.ul_cfg(tbs, qm, 0, nb_q, l_prb, cols, nof_ack, ri_len, cqi, i_ack=int(rng.integers(0, 10)),
                      i_ri=int(rng.integers(0, 12)), i_cqi=int(rng.integers(2, 12)))
        q_ack, q_ri, q_cqi = ol.port_uci_q_primes(u)
        G = nb_q - (q_ri + q_cqi) * qm
        seg = ol.PortCbsegm()
        assert P.port_cbsegm(C.byref(seg), tbs) == 0 and seg.F == 0
        payload = rng.integers(0, 2, tbs, dtype=np.uint8)
        tb = vec.attach_crc(vec.CRC24A, payload[None, :])[0]
        e_parts, pos = [], 0
        Gp, gamma = G // qm, (G // qm) % seg.C
        for cb in range(seg.C):
            K = seg.K1 if cb < seg.C1 else seg.K2
            rlen = K if seg.C == 1 else K - 24
            blk = tb[pos:pos + rlen]
            pos += rlen
            if seg.C > 1:
                blk = vec.attach_crc(vec.CRC24B, blk[None, :])[0]
            E = qm * (Gp // seg.C) if cb <= seg.C - gamma - 1 else qm * ((Gp + seg.C - 1) // seg.C)
            e_parts.append(vec.rate_match(vec.turbo_encode(blk[None, :]), E, 0)[0])
        g_tx = np.concatenate([rng.integers(0, 2, q_cqi * qm, dtype=np.uint8)] + e_parts).astype(np.uint8)
        assert g_tx.size == nb_q - q_ri * qm
        ri_pos = _uci_positions(True, q_ri, qm, rows, cols).reshape(-1)
        ack_pos = _uci_positions(False, q_ack, qm, rows, cols).reshape(-1)
        is_ri = np.zeros(nb_q, bool)
        is_ri[ri_pos] = True
        jj, ii, kk = np.meshgrid(np.arange(rows), np.arange(cols), np.arange(qm), indexing="ij")
        x = (jj * qm + ii * rows * qm + kk).reshape(-1)   # channel positions in UL-SCH (row-major) order
        x = x[~is_ri[x]]
        q_tx = rng.integers(0, 2, nb_q, dtype=np.uint8)     # RI cells: anything
        q_tx[x] = g_tx
        q_tx[ack_pos] = rng.integers(0, 2, ack_pos.size, dtype=np.uint8)
        c_init = int(rng.integers(1, 2 ** 31 - 1))
        c = np.zeros(nb_q, np.uint8)
        P.port_gold_sequence(c_init, nb_q, c)
        sigma = (0.0, 0.07, 0.2)[i % 3]
        sym = _modulate(q_tx ^ c, qm)
        sym = (sym + sigma * (rng.standard_normal(sym.size) + 1j * rng.standard_normal(sym.size))).astype(np.complex64)
        q = ol.port_demod_descramble(qm, sym, c_init, nb_q)
        rc, out, avg, _, qp = ol.port_ulsch_decode(u, q, c, 8)
        assert qp == (q_ack, q_ri, q_cqi)
        want.append((rc, out, avg, np.packbits(payload)))
        descs.append(dict(tbs=tbs, qm=qm, rv=0, nof_e_bits=nb_q, softbuffer=i, c_init=c_init, symbols=sym, ul_nof_symb=cols,
                          uci=dict(q_prime_ack=q_ack, q_prime_ri=q_ri, q_prime_cqi=q_cqi, ri_len=ri_len)))
    pool = ctx.harq_pool(len(descs), 13)
    got = ctx.decode_tb_sym_batch(pool, descs, 8)
    n_ok = 0
    for i, ((ret, data, avg), (rc, out, wavg, payload)) in enumerate(zip(got, want)):
        tbs = descs[i]["tbs"]
        assert ret == rc, i
        assert abs(avg - wavg) < 1e-6, (i, avg, wavg)
        assert np.array_equal(data[: tbs // 8 + 3], out), i
        if ret == 0:
            n_ok += 1
            assert np.array_equal(data[: tbs // 8], payload)
    assert n_ok >= 12
    pool.close()
