"""Parity of the CUDA path (through the C ABI) against the oracle and the golden vectors.  B200 only.

Bar: bit-exact decoded bytes, CRC pass/fail and half-iteration counts -- integer arithmetic, no tolerance.
"""
import numpy as np
import pytest

import oracle_libs as ol

pytestmark = pytest.mark.gpu


def _cases(g, prefix):
    return sorted({k.split("_")[0] for k in g if k.startswith(prefix)}, key=lambda s: int(s[1:]))


def _stop_rule(llr1, K, max_it, poly):
    """sch.c:353-383 applied to the oracle's per-iteration decisions."""
    by, _, _ = ol.port_trace(llr1, K, max_it)
    P = ol.port()
    for it in range(max_it):
        if P.port_crc_bytes(poly, by[it].copy(), K) == 0:
            return it + 1, 1, by[it]
    return max_it, 0, by[max_it - 1]


def test_golden_vectors_every_iteration(ctx, golden):
    g = golden["tdec_vectors"]
    for c in _cases(g, "c"):
        K = int(g[f"{c}_K"][0])
        llr, dec = g[f"{c}_llr"], g[f"{c}_dec"]
        for nit in range(1, 11):
            got, n_iter, _ = ctx.tdec_batch_host(llr, K, nit)
            assert np.array_equal(got, dec[:, nit - 1]), (c, K, nit)
            assert (n_iter == nit).all()


@pytest.mark.parametrize("K", [40, 48, 104, 400, 408, 512, 800, 816, 1024, 2048, 5824, 6144])
def test_noisy_blocks_vs_oracle(ctx, vec, K):
    for si, (sigma, scale) in enumerate([(1.457, 100), (1.092, 100), (0.9, 700), (0.6, 4000)]):
        bits, llr = vec.make_blocks(5, K, sigma, scale, seed=10 * K + si)
        for nit in (0, 1, 2, 3, 4, 8):
            got, n_iter, ok = ctx.tdec_batch_host(llr, K, nit)
            assert np.array_equal(got, ol.port_run_all(llr, K, nit)), (K, sigma, scale, nit)
            assert (n_iter == max(nit, 1)).all() and (ok == 0).all()


def test_all_188_sizes_mixed_batch(ctx, vec):
    """BASELINE config 4: every LTE code-block size in ONE mixed-K batch (per-block window sizing)."""
    Ks, llrs = [], []
    stride = 3 * 6144 + 12
    for K in ol.ALL_K:
        bits, llr = vec.make_blocks(2, K, 1.092, 100, seed=K)
        pad = np.zeros((2, stride), np.int16)
        pad[:, : llr.shape[1]] = llr
        llrs.append(pad)
        Ks += [K, K]
    llr = np.ascontiguousarray(np.concatenate(llrs))
    Ks = np.array(Ks, np.uint32)
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(Ks))          # interleave sizes so the scheduler has to regroup them
    llr, Ks = np.ascontiguousarray(llr[perm]), Ks[perm]
    for nit in (1, 4, 5):
        got, n_iter, _ = ctx.tdec_batch_host(llr, Ks, nit)
        for i, K in enumerate(Ks):
            K = int(K)
            want = ol.port_run_all(llr[i:i + 1, : 3 * K + 12].copy(), K, nit)[0]
            assert np.array_equal(got[i, : K // 8], want), (K, nit)
        assert (n_iter == nit).all()


@pytest.mark.parametrize("K", [104, 512, 1024, 6144])
def test_crc_early_termination(ctx, vec, pkg, K):
    """BASELINE config 3 semantics: CRC after every half iteration, frozen output, per-block counts."""
    bits, llr = vec.make_blocks(16, K, vec.harness_sigma(4.0), 100, seed=K)
    llr[0] = vec.awgn_llr(vec.turbo_encode(bits[:1]), 0.0)[0]            # a clean block: stops after 1
    llr[1] = vec.make_blocks(1, K, 3.0, 100, seed=1, crc=False)[1][0]    # garbage: runs to the cap
    for max_it in (1, 4, 10):
        got, n_iter, ok = ctx.tdec_batch_host(llr, K, max_it, crc_mode=pkg.CRC_24B)
        for i in range(llr.shape[0]):
            n, o, by = _stop_rule(llr[i], K, max_it, ol.CRC24B)
            assert (int(n_iter[i]), int(ok[i])) == (n, o), (K, max_it, i)
            assert np.array_equal(got[i], by), (K, max_it, i)
        assert n_iter[0] == 1 and ok[0] == 1 and ok[1] == 0
    # CRC24A variant (single-code-block transport block)
    payload = np.random.default_rng(K).integers(0, 2, (4, K - 24), dtype=np.uint8)
    b = vec.attach_crc(vec.CRC24A, payload)
    llr = vec.awgn_llr(vec.turbo_encode(b), 0.8, 100, np.random.default_rng(1))
    got, n_iter, ok = ctx.tdec_batch_host(llr, K, 8, crc_mode=pkg.CRC_24A)
    for i in range(4):
        n, o, by = _stop_rule(llr[i], K, 8, ol.CRC24A)
        assert (int(n_iter[i]), int(ok[i])) == (n, o) and np.array_equal(got[i], by)


def test_working_layout_input(ctx, vec):
    """input as srslte_rm_turbo_rx_lut leaves it in the soft buffer (sub-block layout), sch.c style."""
    for K in (408, 800, 816, 6144, 40):
        bits, llr = vec.make_blocks(4, K, 1.092, 100, seed=K)
        sb = vec.sb_layout_from_natural(llr, K)
        stride = (sb.shape[1] + 1) // 2 * 2
        buf = np.zeros((4, stride), np.int16)
        buf[:, : sb.shape[1]] = sb
        got, _, _ = ctx.tdec_batch_host(buf, K, 5, natural=False)
        assert np.array_equal(got, ol.port_run_all(sb, K, 5, natural=False)), K


def test_edge_cases(ctx, vec, pkg):
    # empty batch
    got, n_iter, ok = ctx.tdec_batch_host(np.zeros((0, 3 * 40 + 12), np.int16), 40, 4)
    assert got.shape == (0, 5)
    # invalid K, odd stride, short stride -> SRSLTE_ERROR_INVALID_INPUTS, like the reference's -2
    for bad in (41, 6208, 0):
        with pytest.raises(pkg.B200Error, match="-2"):
            ctx.tdec_batch_host(np.zeros((1, 3 * 6208 + 12), np.int16), bad, 1)
    with pytest.raises(pkg.B200Error, match="-2"):
        ctx.tdec_batch_host(np.zeros((1, 3 * 40 + 10), np.int16), 40, 1)
    # all-zero and extreme-value LLRs decode like the oracle (wrap / saturation corners)
    for K in (40, 512, 6144):
        for fill in (0, 32767, -32768):
            llr = np.full((2, 3 * K + 12), fill, np.int16)
            got, _, _ = ctx.tdec_batch_host(llr, K, 3)
            assert np.array_equal(got, ol.port_run_all(llr, K, 3)), (K, fill)
        rng = np.random.default_rng(K)
        llr = rng.integers(-32768, 32768, (3, 3 * K + 12)).astype(np.int16)
        got, _, _ = ctx.tdec_batch_host(llr, K, 4)
        assert np.array_equal(got, ol.port_run_all(llr, K, 4)), K
    # ragged batch sizes around the blocks-per-warp boundaries
    for K, n in ((6144, 1), (6144, 5), (512, 9), (40, 33)):
        bits, llr = vec.make_blocks(n, K, 1.092, 100, seed=n)
        got, _, _ = ctx.tdec_batch_host(llr, K, 2)
        assert np.array_equal(got, ol.port_run_all(llr, K, 2)), (K, n)


def test_full_size_round_trip_properties(ctx, vec, pkg):
    """BASELINE-size batch through size-independent properties: noiseless encode -> decode returns the
    payload after ONE half iteration with CRC ok for every block; the batch result does not depend on
    batch composition (each block decodes as it does alone)."""
    K, n = 6144, 8192
    rng = np.random.default_rng(42)
    payload = rng.integers(0, 2, (64, K - 24), dtype=np.uint8)
    bits = vec.attach_crc(vec.CRC24B, payload)
    coded = vec.turbo_encode(bits)
    idx = np.arange(n) % 64
    llr = np.ascontiguousarray(((coded[idx].astype(np.int16) * 2 - 1) * 100))
    got, n_iter, ok = ctx.tdec_batch_host(llr, K, 8, crc_mode=pkg.CRC_24B)
    assert (n_iter == 1).all() and (ok == 1).all()
    assert np.array_equal(got, np.packbits(bits, axis=1)[idx])
    # noisy: a big batch equals the same blocks decoded in small batches (no cross-block coupling)
    bits, small = vec.make_blocks(24, K, 1.092, 100, seed=3)
    big = np.ascontiguousarray(small[np.arange(1000) % 24])
    a, na, oa = ctx.tdec_batch_host(big, K, 8, crc_mode=pkg.CRC_24B)
    b, nb, ob = ctx.tdec_batch_host(small, K, 8, crc_mode=pkg.CRC_24B)
    assert np.array_equal(a, b[np.arange(1000) % 24]) and np.array_equal(na, nb[np.arange(1000) % 24])
    want = ol.port_run_all(small, K, 4)
    c, _, _ = ctx.tdec_batch_host(big, K, 4)
    assert np.array_equal(c, want[np.arange(1000) % 24])


def test_device_pointer_entry_matches_host_entry(ctx, vec, pkg):
    import torch
    K, n, nit = 6144, 64, 4
    bits, llr = vec.make_blocks(n, K, 1.092, 100, seed=9)
    want, _, _ = ctx.tdec_batch_host(llr, K, nit)
    d = torch.from_numpy(llr).cuda()
    out = torch.zeros((n, K // 8), dtype=torch.uint8, device="cuda")
    nit_d = torch.zeros(n, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ctx.set_stream(s.cuda_stream)
        ctx.tdec_batch_dev(d.data_ptr(), n, llr.shape[1], K, nit, out.data_ptr(), K // 8, nit_d.data_ptr(), 0)
        s.synchronize()
    ctx.set_stream(0)
    assert np.array_equal(out.cpu().numpy(), want) and (nit_d.cpu().numpy() == nit).all()
    assert ctx.launch_count > 0


def test_rate_dematch_vs_oracle(ctx, vec, pkg):
    """srslte_rm_turbo_rx_lut semantics on the GPU: scatter-add with HARQ combining and wrap-around."""
    import torch
    P = ol.port()
    rng = np.random.default_rng(4)
    blocks, e_parts, want = [], [], []
    e_off, stride = 0, 18600
    cases = [(6144, 0, 6918), (6144, 2, 20000), (5824, 1, 3 * 5824 + 12), (816, 3, 5000), (512, 0, 9 * 512),
             (408, 2, 700), (400, 1, 1500), (40, 3, 1000)]
    work0 = rng.integers(-20000, 20000, (len(cases), stride)).astype(np.int16)
    for i, (K, rv, E) in enumerate(cases):
        e = rng.integers(-25000, 25000, E).astype(np.int16)
        blocks.append((K, rv, e_off, E, i * stride))
        e_parts.append(e)
        e_off += E + (E & 1)
        if E & 1:
            e_parts.append(np.zeros(1, np.int16))
        w = work0[i].copy()
        assert P.port_rm_turbo_rx(e, E, w, K, rv, 1) == 0
        want.append(w)
    e_all = torch.from_numpy(np.concatenate(e_parts)).cuda()
    work = torch.from_numpy(work0.copy()).cuda()
    ctx.rm_rx_batch_dev(blocks, e_all.data_ptr(), work.data_ptr())
    ctx.synchronize()
    torch.cuda.synchronize()
    assert np.array_equal(work.cpu().numpy(), np.stack(want))
    # and the de-matched buffer decodes: rate match -> dematch on GPU -> decode (working layout) -> payload
    K, E, rv = 6144, 9000, 0
    bits, _ = vec.make_blocks(4, K, 0.0, seed=1)
    e = ((vec.rate_match(vec.turbo_encode(bits), E, rv).astype(np.int16) * 2 - 1) * 100)
    e_d = torch.from_numpy(np.ascontiguousarray(e)).cuda()
    wl = (pkg.working_len(K) + 63) // 64 * 64
    work = torch.zeros((4, wl), dtype=torch.int16, device="cuda")
    ctx.rm_rx_batch_dev([(K, rv, i * E, E, i * wl) for i in range(4)], e_d.data_ptr(), work.data_ptr())
    ctx.synchronize()
    got, n_iter, ok = ctx.tdec_batch_host(work.cpu().numpy(), K, 8, crc_mode=pkg.CRC_24B, natural=False)
    assert np.array_equal(got, np.packbits(bits, axis=1)) and (ok == 1).all()


def test_fast_and_exact_variants_agree(ctx, vec, pkg):
    """The fast (wrapping, proven) and the exact (saturating) variants of the window decoders give the
    same bytes; benign inputs stay on the fast variant, saturating inputs fall back to the exact one."""
    for K in (408, 1024, 6144):
        for sigma, scale in ((1.092, 100), (0.9, 700), (0.6, 4000)):
            bits, llr = vec.make_blocks(8, K, sigma, scale, seed=K + scale)
            want = ol.port_run_all(llr, K, 6)
            f0 = ctx.fallback_count
            fast, _, _ = ctx.tdec_batch_host(llr, K, 6)
            f1 = ctx.fallback_count
            ctx.set_exact(True)
            exact, _, _ = ctx.tdec_batch_host(llr, K, 6)
            ctx.set_exact(False)
            assert np.array_equal(fast, want) and np.array_equal(exact, want), (K, sigma, scale)
            if scale == 4000:
                assert f1 > f0, "inputs that saturate must take the exact re-run"
    # the benchmark operating point never needs the exact re-run
    bits, llr = vec.make_blocks(64, 6144, vec.harness_sigma(1.5), 100, seed=77, crc=False)
    f0 = ctx.fallback_count
    got, _, _ = ctx.tdec_batch_host(llr, 6144, 4)
    assert ctx.fallback_count == f0
    assert np.array_equal(got, ol.port_run_all(llr, 6144, 4))


def test_near_saturation_boundary(ctx, vec):
    """LLR scales swept across the point where the reference's saturating adds start to clamp: the
    proof-or-fallback logic must stay bit-exact on both sides of it."""
    K = 2048
    for scale in (150, 300, 450, 600, 900, 1300, 2000):
        bits, llr = vec.make_blocks(4, K, 0.8, scale, seed=scale)
        for nit in (2, 5, 9):
            got, _, _ = ctx.tdec_batch_host(llr, K, nit)
            assert np.array_equal(got, ol.port_run_all(llr, K, nit)), (scale, nit)


def test_pure_fast_tier_edges(ctx, vec):
    """The tier without exact edge rows (G <= 2529, L % 4 == 0) next to its limits: inputs whose bound G sits just
    below / above the tier thresholds, adversarial sign patterns in the rows next to the known start state and
    in the tail, and tail samples that dominate the bound.  Bit-exact with the oracle in every case."""
    rng = np.random.default_rng(2529)
    for K in (6144, 1024, 512):
        base_bits, base = vec.make_blocks(6, K, 0.9, 100, seed=K + 5)
        cases = []
        # (a) body magnitudes clipped so that max|sys| + max|par| is just below / above 2529 and 2978
        for lim in (1200, 1264, 1265, 1400, 1489, 1490, 1600):
            x = np.clip(base.astype(np.int32) * 12, -lim, lim).astype(np.int16)
            cases.append(("clip%d" % lim, x))
        # (b) constant-magnitude inputs with random / all-equal signs: the worst case of the bound on every row
        for mag in (1260, 1264):
            s = rng.integers(0, 2, base.shape) * 2 - 1
            cases.append(("pm%d" % mag, (s * mag).astype(np.int16)))
            cases.append(("pos%d" % mag, np.full(base.shape, mag, np.int16)))
            cases.append(("neg%d" % mag, np.full(base.shape, -mag, np.int16)))
        # (c) small body, huge tail samples (the tail must count for G), and the other way round
        x = (base // 4).astype(np.int16)
        x[:, 3 * K:] = (rng.integers(0, 2, (base.shape[0], 12)) * 2 - 1) * 9000
        cases.append(("bigtail", x))
        x = np.clip(base.astype(np.int32) * 12, -1264, 1264).astype(np.int16)
        x[:, 3 * K:] = 0
        x[:, :24] = 1264  # rows 0..7 of window 0 all at the bound
        cases.append(("start", x))
        for name, llr in cases:
            for nit in (1, 2, 4):
                got, _, _ = ctx.tdec_batch_host(llr, K, nit)
                assert np.array_equal(got, ol.port_run_all(llr, K, nit)), (K, name, nit)


def test_generic_decoder_pairs_and_ragged_items(ctx, vec, pkg):
    """K <= 400: a thread decodes two blocks and a warp item holds up to 64 -- odd counts, single blocks, several
    items, with and without CRC early termination (n_iter and the CRC flag per block), extreme LLR values."""
    P = ol.port()
    for K in (40, 104, 400):
        for n in (1, 2, 3, 63, 64, 65, 131):
            bits, llr = vec.make_blocks(n, K, vec.harness_sigma(4.0), 100, seed=K * 7 + n)
            got, _, _ = ctx.tdec_batch_host(llr, K, 5)
            assert np.array_equal(got, ol.port_run_all(llr, K, 5)), (K, n)
        bits, llr = vec.make_blocks(67, K, 0.7, 100, seed=K + 3)
        llr[5] = 32767
        llr[6] = -32768
        got, n_iter, ok = ctx.tdec_batch_host(llr, K, 8, crc_mode=pkg.CRC_24B)
        for i in range(67):
            by, _, _ = ol.port_trace(llr[i], K, 8)
            crcs = [P.port_crc_bytes(ol.CRC24B, by[it].copy(), K) for it in range(8)]
            stop = next((it + 1 for it in range(8) if crcs[it] == 0), 8)
            assert n_iter[i] == stop and ok[i] == int(crcs[stop - 1] == 0), (K, i)
            assert np.array_equal(got[i], by[stop - 1]), (K, i)
