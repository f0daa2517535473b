"""Block-granular early termination (tdec_win_dyn_kernel) and the single-process multi-device entry.  B200 only.

reference behaviour: lib/src/phy/phch/sch.c:353-383 -- a code block stops at the half iteration its CRC passes; the
bytes, the half-iteration count and the CRC flag of every block must not depend on which blocks shared a warp with
it, in which order the thread groups picked their blocks up, or which kernel ran them.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_libs as ol

pytestmark = pytest.mark.gpu

ROUND_BASED = 64   # srslte_b200_ctx_set_variant_bits: CRC modes through the round-based kernel


def _mixed_convergence(vec, n, K, seed):
    """blocks whose convergence time differs: noise levels from hopeless to easy, shuffled"""
    rng = np.random.default_rng(seed)
    parts = []
    for i, e_db in enumerate((1.5, 3.2, 4.0, 4.0, 5.0, 7.0)):
        _, l = vec.make_blocks((n + 5) // 6, K, vec.harness_sigma(e_db), 100.0, seed=seed + i)
        parts.append(l)
    llr = np.concatenate(parts)[:n]
    return np.ascontiguousarray(llr[rng.permutation(n)])


def _both_kernels(ctx, llr, K, nit, mode):
    dyn = ctx.tdec_batch_host(llr, K, nit, crc_mode=mode)
    ctx.set_variant_bits(ROUND_BASED)
    try:
        rnd = ctx.tdec_batch_host(llr, K, nit, crc_mode=mode)
    finally:
        ctx.set_variant_bits(0)
    return dyn, rnd


@pytest.mark.parametrize("K,n", [(6144, 4099), (2112, 2501), (1024, 3001), (512, 3003), (1008, 1501), (504, 1203)])
def test_dyn_kernel_equals_round_based_kernel(ctx, pkg, vec, K, n):
    """W = 16 and W = 8, L % 16 = 0 / 4 (main path) and L % 4 != 0 (general path: warps refilled as a whole); n leaves a
    partial last item.  Bytes, half-iteration counts and flags of both kernels agree, and a sample agrees with the oracle."""
    llr = _mixed_convergence(vec, n, K, seed=K)
    for nit in (8, 3, 1, 0):
        dyn, rnd = _both_kernels(ctx, llr, K, nit, pkg.CRC_24B)
        assert np.array_equal(dyn[1], rnd[1]), np.nonzero(dyn[1] != rnd[1])[0][:8]
        assert np.array_equal(dyn[2], rnd[2])
        assert np.array_equal(dyn[0], rnd[0]), np.nonzero((dyn[0] != rnd[0]).any(axis=1))[0][:8]
        if nit >= 3:
            assert 0 < int(dyn[2].sum()) < n            # some converge, some do not
            assert len(np.unique(dyn[1])) >= (3 if nit > 3 else 2)   # at different half iterations
        else:
            assert (dyn[1] == 1).all()                  # nof_iterations 0 runs one half iteration, like srslte_tdec_run_all
    P = ol.port()
    for i in range(0, n, max(1, n // 6)):
        by, _, _ = ol.port_trace(llr[i], K, 8)
        crcs = [P.port_crc_bytes(ol.CRC24B, by[it].copy(), K) for it in range(8)]
        stop = next((it + 1 for it in range(8) if crcs[it] == 0), 8)
        d = ctx.tdec_batch_host(llr[i:i + 1], K, 8, crc_mode=pkg.CRC_24B)
        full = ctx.tdec_batch_host(llr, K, 8, crc_mode=pkg.CRC_24B)
        for got in ((d[0][0], d[1][0], d[2][0]), (full[0][i], full[1][i], full[2][i])):
            assert int(got[1]) == stop and int(got[2]) == int(crcs[stop - 1] == 0), (K, i)
            assert np.array_equal(got[0], by[stop - 1]), (K, i)


def test_dyn_kernel_mixed_sizes_and_per_block_crc_modes(ctx, pkg, vec):
    """several block sizes (epochs) in one launch, through the transport-block entry's per-block CRC modes is covered by
    test_gpu_transport_block; here a mixed-K batch with CRC24A"""
    L = pkg.lib()
    sizes = [6144, 5824, 3136, 1024, 1008, 816, 512, 408, 200, 40]
    per = 37
    Ks = np.repeat(np.array(sizes, dtype=np.uint32), per)
    rng = np.random.default_rng(5)
    Ks = Ks[rng.permutation(len(Ks))]
    stride = 3 * 6144 + 12
    llr = np.zeros((len(Ks), stride), np.int16)
    for K in sizes:
        rows = np.nonzero(Ks == K)[0]
        l = _mixed_convergence(vec, len(rows), K, seed=11 * K)
        llr[rows, : 3 * K + 12] = l
    res = []
    for bits in (0, ROUND_BASED):
        ctx.set_variant_bits(bits)
        try:
            out = np.zeros((len(Ks), 768), np.uint8)
            nout = np.zeros(len(Ks), np.uint8)
            okout = np.zeros(len(Ks), np.uint8)
            b = pkg.TdecBatch()
            b.n_cb = len(Ks); b.long_cb = Ks.ctypes.data_as(C.POINTER(C.c_uint32)); b.uniform_long_cb = 0
            b.in_stride = stride; b.out_stride = 768; b.nof_iterations = 9; b.crc_mode = pkg.CRC_24B; b.input_format = 0
            assert L.srslte_b200_tdec_batch_host(ctx._h, C.byref(b), llr.ctypes.data_as(C.c_void_p),
                                                 out.ctypes.data_as(C.c_void_p), nout.ctypes.data_as(C.c_void_p),
                                                 okout.ctypes.data_as(C.c_void_p)) == 0
            res.append((out, nout, okout))
        finally:
            ctx.set_variant_bits(0)
    for a, b_ in zip(res[0], res[1]):
        assert np.array_equal(a, b_)
    assert 0 < int(res[0][2].sum()) < len(Ks)


def test_dyn_kernel_forced_tiers(ctx, pkg, vec):
    """static / tracked tier forced, general path only, refill at once (parities mix inside a warp), exact variant forced:
    same results"""
    K, n = 1024, 1999
    llr = _mixed_convergence(vec, n, K, seed=99)
    want = ctx.tdec_batch_host(llr, K, 8, crc_mode=pkg.CRC_24B)
    for bits in (2, 6, 8, 512):
        ctx.set_variant_bits(bits)
        try:
            got = ctx.tdec_batch_host(llr, K, 8, crc_mode=pkg.CRC_24B)
        finally:
            ctx.set_variant_bits(0)
        for a, b_ in zip(got, want):
            assert np.array_equal(a, b_), bits
    ctx.set_exact(True)
    try:
        got = ctx.tdec_batch_host(llr, K, 8, crc_mode=pkg.CRC_24B)
    finally:
        ctx.set_exact(False)
    for a, b_ in zip(got, want):
        assert np.array_equal(a, b_)


@pytest.mark.parametrize("scale", [100.0, 150.0, 230.0])
def test_dyn_kernel_at_the_config3_operating_point(ctx, pkg, vec, scale):
    """K = 6144 at harness -e 4.0 (BASELINE config 3's converging case): blocks converge at different half iterations, their
    extrinsic values grow and the static / tracked tiers (larger LLR scales: also the exact fallback) mix with the pure one
    inside a warp whose groups sit at different parities.  Every output of the block-granular kernel -- with the tracked
    tier through the main path (default) and with everything through the general path (bit 3: warps are then refilled
    as a whole) -- equals the round-based kernel's."""
    K, n = 6144, 6000
    _, llr = vec.make_blocks(n, K, vec.harness_sigma(4.0), scale, seed=int(scale))
    t0 = ctx.tier_counts
    dyn, rnd = _both_kernels(ctx, llr, K, 8, pkg.CRC_24B)
    tiers = [a - b for a, b in zip(ctx.tier_counts, t0)]
    ctx.set_variant_bits(8)
    try:
        gen = ctx.tdec_batch_host(llr, K, 8, crc_mode=pkg.CRC_24B)
    finally:
        ctx.set_variant_bits(0)
    for name, got in (("main", dyn), ("general", gen)):
        assert np.array_equal(got[1], rnd[1]), (name, np.nonzero(got[1] != rnd[1])[0][:8])
        assert np.array_equal(got[2], rnd[2]), name
        assert np.array_equal(got[0], rnd[0]), (name, np.nonzero((got[0] != rnd[0]).any(axis=1))[0][:8])
    assert tiers[1] > 0 and tiers[2] > 0, tiers      # static and tracked tiers ran
    assert 0.5 < rnd[2].mean() < 1.0


def test_early_termination_buys_time(ctx, pkg, vec):
    """a batch in which most blocks converge after a few half iterations must be much cheaper than one that never
    converges (the round-based kernel needed the full time for both)"""
    import torch
    K, n = 6144, 148 * 32 * 2
    t = {}
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    for name, e_db in (("never", 1.0), ("fast", 6.0)):
        _, llr = vec.make_blocks(512, K, vec.harness_sigma(e_db), 100.0, seed=3)
        llr = np.ascontiguousarray(np.tile(llr, ((n + 511) // 512, 1))[:n])
        d_llr = torch.from_numpy(llr).cuda()
        d_out = torch.zeros((n, K // 8), dtype=torch.uint8, device="cuda")
        d_nit = torch.zeros(n, dtype=torch.uint8, device="cuda")
        d_ok = torch.zeros(n, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for rep in range(3):
            if rep == 1:
                ev[0].record(stream)
            ctx.tdec_batch_dev(d_llr.data_ptr(), n, llr.shape[1], K, 8, d_out.data_ptr(), K // 8, d_nit.data_ptr(),
                               d_ok.data_ptr(), crc_mode=pkg.CRC_24B)
        ev[1].record(stream)
        torch.cuda.synchronize()
        t[name] = ev[0].elapsed_time(ev[1]) / 2
        t[name + "_its"] = float(d_nit.float().mean())
    assert t["never_its"] == 8.0 and t["fast_its"] < 4.0, t
    # ideal: fast / never = its ratio; allow the layout pass and the tail
    torch.cuda.set_stream(torch.cuda.default_stream())
    assert t["fast"] < t["never"] * (t["fast_its"] / 8.0 + 0.3), t


def test_group_entry_single_process(pkg, vec):
    """srslte_b200_group_*: the same results as one context, whatever the number of devices (1 on the round-end box)"""
    import torch
    ndev = torch.cuda.device_count()
    K, n = 2048, 1001
    _, llr = vec.make_blocks(n, K, vec.harness_sigma(3.0), 100.0, seed=21)
    one = pkg.Context(0)
    want = one.tdec_batch_host(llr, K, 6, crc_mode=pkg.CRC_24B)
    one.close()
    for nd in sorted({1, ndev, min(2, ndev)}):
        g = pkg.Group(nd)
        assert len(g) == nd
        got = g.tdec_batch_host(llr, K, 6, crc_mode=pkg.CRC_24B)
        for a, b_ in zip(got, want):
            assert np.array_equal(a, b_), nd
        g.set_weights([1.0 + 2.0 * i for i in range(nd)])       # unequal shares: same results
        got = g.tdec_batch_host(llr, K, 6, crc_mode=pkg.CRC_24B)
        for a, b_ in zip(got, want):
            assert np.array_equal(a, b_), nd
        assert len(g.calibrate()) == nd
        got = g.tdec_batch_host(llr, K, 6, crc_mode=pkg.CRC_24B)
        for a, b_ in zip(got, want):
            assert np.array_equal(a, b_), nd
        pin = pkg.PinnedArray((nd, 8 << 20), np.uint8)
        gbs = g.h2d_probe(pin.array.ctypes.data, 8 << 20, reps=4)
        assert len(gbs) == nd and all(x > 1.0 for x in gbs), gbs
        pin.free()
        g.close()


@pytest.mark.parametrize("K,n", [(6144, 261), (2112, 300), (512, 203), (1008, 77)])
def test_plain_batches_kernel_without_crc_variants(ctx, pkg, vec, K, n):
    """launches without CRC run the kernel that carries no CRC variants (default) or the combined one (variant bit 7):
    same bytes, and the oracle's on a sample"""
    llr = _mixed_convergence(vec, n, K, seed=3 * K)
    llr[: n // 3] = (llr[: n // 3].astype(np.int32) * 30).clip(-32768, 32767).astype(np.int16)   # static / tracked / exact tiers
    for nit in (1, 4, 7):
        a = ctx.tdec_batch_host(llr, K, nit)
        ctx.set_variant_bits(128)
        try:
            b = ctx.tdec_batch_host(llr, K, nit)
        finally:
            ctx.set_variant_bits(0)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    idx = np.array([0, 1, n // 3 + 1, n - 1])
    assert np.array_equal(a[0][idx], ol.port_run_all(np.ascontiguousarray(llr[idx]), K, 7))
