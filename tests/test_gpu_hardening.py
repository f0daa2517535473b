"""Parity hardening (VERDICT r01 "Next round" item 6).  B200 only.

1. Differential fuzz ON THE GPU: every block decoded by the fast tiers (wrapping arithmetic + proofs) and again with
   the exact saturating variant forced (srslte_b200_ctx_set_exact), which is pinned to the oracle elsewhere.  >= 100 000
   blocks per window count, LLR magnitudes swept so that the pure, static, tracked tiers and the exact fallback all run;
   plus single-half-iteration blocks whose bound G sits exactly at 2529 / 2978 / 5461 +- 3.
2. BASELINE config 4 at the survey's spec: all 188 sizes x 64 blocks, nof_iterations 1..10, two noise levels, against the
   compiled reference when it is there (else the port), and CRC mode with iteration counts and flags on a subset.
3. BASELINE config 5 style HARQ: 200 transport blocks at LLR scale 400 with the rv sequence 0, 2, 3, 1 on the same soft
   buffers, every transmission against the oracle's port_decode_tb.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_libs as ol

pytestmark = pytest.mark.gpu


def _decode_both(ctx, llr, K, nit):
    got, _, _ = ctx.tdec_batch_host(llr, K, nit)
    ctx.set_exact(True)
    try:
        want, _, _ = ctx.tdec_batch_host(llr, K, nit)
    finally:
        ctx.set_exact(False)
    return got, want


@pytest.mark.parametrize("K", [1024, 512, 2112, 6144])       # W=16 (L%16=0), W=8, W=16 with bottom rows, the headline size
def test_fast_tiers_equal_exact_variant_fuzz(ctx, K):
    rng = np.random.default_rng(K)
    n_total = {1024: 104_000, 512: 104_000, 2112: 24_000, 6144: 12_000}[K]
    t0 = ctx.tier_counts
    done = 0
    # amplitude sets the tier: (sys, parity) uniform in +-amp; with nit >= 2 the extrinsic values move G further
    for amp, nit in ((600, 4), (1250, 1), (1250, 3), (1400, 1), (1480, 2), (2000, 1), (2600, 2), (2700, 1), (9000, 2), (30000, 3)):
        n = n_total // 10
        llr = rng.integers(-amp, amp + 1, (n, 3 * K + 12), dtype=np.int64).astype(np.int16)
        got, want = _decode_both(ctx, llr, K, nit)
        bad = np.nonzero((got != want).any(axis=1))[0]
        assert bad.size == 0, (K, amp, nit, bad[:5])
        done += n
    tiers = [a - b for a, b in zip(ctx.tier_counts, t0)]
    assert done >= 12_000
    assert all(t > 0 for t in tiers), tiers        # pure, static, tracked and exact variants all ran


@pytest.mark.parametrize("K", [1024, 512, 6144])
def test_tier_thresholds_exactly(ctx, K):
    """one half iteration, G = max|sys| + max|par0| planted exactly at the tier thresholds +- 3"""
    rng = np.random.default_rng(7 * K)
    blocks = []
    for thr in (2529, 2978, 5461):
        for dG in range(-3, 4):
            G = thr + dG
            for split in (0.3, 0.5, 0.7):
                smax = int(G * split)
                pmax = G - smax
                b = np.zeros(3 * K + 12, np.int16)
                b[0::3][:K] = rng.integers(-smax, smax + 1, K)
                b[1::3][:K] = rng.integers(-pmax, pmax + 1, K)
                b[2::3][:K] = rng.integers(-pmax, pmax + 1, K)
                b[3 * K:] = rng.integers(-min(smax, pmax), min(smax, pmax) + 1, 12)
                # plant the maxima (several positions incl. the first rows next to the known start state)
                for pos in (0, 3, 3 * (K // 2), 3 * (K - 1)):
                    b[pos] = smax if (pos // 3) % 2 == 0 else -smax
                    b[pos + 1] = -pmax if (pos // 3) % 2 == 0 else pmax
                blocks.append(b)
    llr = np.stack(blocks)
    for nit in (1, 2):
        got, want = _decode_both(ctx, llr, K, nit)
        assert np.array_equal(got, want), (K, nit, np.nonzero((got != want).any(axis=1))[0][:5])
        assert np.array_equal(got, ol.port_run_all(llr, K, nit))      # and the oracle itself on these few blocks


def test_config4_all_sizes_at_spec(ctx, pkg, vec):
    """all 188 sizes x 64 blocks in ONE mixed batch, nof_iterations 1..10, harness -e 1.5 and -e 4.0"""
    L = pkg.lib()
    per = 64
    Ks = np.repeat(np.array(vec.ALL_K, dtype=np.uint32), per)
    stride = 3 * 6144 + 12
    use_ref = ol.ref() is not None
    for e_db in (1.5, 4.0):
        llr = np.zeros((len(Ks), stride), np.int16)
        for K in vec.ALL_K:
            rows = np.nonzero(Ks == K)[0]
            _, l = vec.make_blocks(per, K, vec.harness_sigma(e_db), 100.0, seed=K + int(10 * e_db), crc=False)
            llr[rows[0]:rows[-1] + 1, : 3 * K + 12] = l
        for nit in range(1, 11):
            out = np.zeros((len(Ks), 768), np.uint8)
            nout = np.zeros(len(Ks), np.uint8)
            b = pkg.TdecBatch()
            b.n_cb = len(Ks); b.long_cb = Ks.ctypes.data_as(C.POINTER(C.c_uint32)); b.uniform_long_cb = 0
            b.in_stride = stride; b.out_stride = 768; b.nof_iterations = nit; b.crc_mode = pkg.CRC_NONE; b.input_format = 0
            rc = L.srslte_b200_tdec_batch_host(ctx._h, C.byref(b), llr.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                               nout.ctypes.data_as(C.c_void_p), C.c_void_p(0))
            assert rc == 0
            assert (nout == nit).all()
            for K in vec.ALL_K:
                rows = np.nonzero(Ks == K)[0]
                blk = np.ascontiguousarray(llr[rows[0]:rows[-1] + 1, : 3 * K + 12])
                want = ol.ref_run_all(blk, K, nit) if use_ref else (ol.port_run_all(blk[:4], K, nit) if nit in (1, 4, 9) else None)
                if want is not None:
                    assert np.array_equal(out[rows[0]:rows[0] + len(want), : K // 8], want), (e_db, nit, K)


def test_config4_crc_mode_counts_and_flags(ctx, pkg, vec):
    """CRC24B early termination over all sizes (4 blocks each): bytes, half-iteration counts and flags vs the oracle"""
    P = ol.port()
    L = pkg.lib()
    per = 4
    sizes = [K for K in vec.ALL_K if K > 40]
    Ks = np.repeat(np.array(sizes, dtype=np.uint32), per)
    stride = 3 * 6144 + 12
    llr = np.zeros((len(Ks), stride), np.int16)
    for K in sizes:
        rows = np.nonzero(Ks == K)[0]
        _, l = vec.make_blocks(per, K, vec.harness_sigma(4.0 if K > 1000 else 5.5), 100.0, seed=3 * K)
        llr[rows[0]:rows[-1] + 1, : 3 * K + 12] = l
    out = np.zeros((len(Ks), 768), np.uint8)
    nout = np.zeros(len(Ks), np.uint8)
    okout = np.zeros(len(Ks), np.uint8)
    b = pkg.TdecBatch()
    b.n_cb = len(Ks); b.long_cb = Ks.ctypes.data_as(C.POINTER(C.c_uint32)); b.uniform_long_cb = 0
    b.in_stride = stride; b.out_stride = 768; b.nof_iterations = 10; b.crc_mode = pkg.CRC_24B; b.input_format = 0
    assert L.srslte_b200_tdec_batch_host(ctx._h, C.byref(b), llr.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                         nout.ctypes.data_as(C.c_void_p), okout.ctypes.data_as(C.c_void_p)) == 0
    passed = 0
    for i in range(0, len(Ks), 2):      # every second block
        K = int(Ks[i])
        by, _, _ = ol.port_trace(llr[i, : 3 * K + 12], K, 10)
        crcs = [P.port_crc_bytes(ol.CRC24B, by[it].copy(), K) for it in range(10)]
        stop = next((it + 1 for it in range(10) if crcs[it] == 0), 10)
        assert int(nout[i]) == stop and int(okout[i]) == int(crcs[stop - 1] == 0), (K, i)
        assert np.array_equal(out[i, : K // 8], by[stop - 1]), (K, i)
        passed += int(okout[i])
    assert passed > len(Ks) // 8       # the operating point makes a good part of the blocks converge


def test_config5_harq_rv_mix_vs_oracle(ctx, vec):
    """200 TBs (16QAM / QPSK / 64QAM, mixed sizes) per call at LLR scale 400, transmissions rv 0, 2, 3, 1 into the same
    soft buffers: return code, bytes, iteration average and per-block CRC flags of EVERY transmission vs port_decode_tb"""
    P = ol.port()
    rng = np.random.default_rng(55)
    sizes = [(2216, 4, 4800), (6200, 4, 9600), (14112, 4, 28800), (4584, 4, 7200), (1000, 2, 2400), (36696, 6, 60000),
             (3624, 4, 5760), (20616, 4, 36000), (9144, 4, 14400), (55056, 6, 72000)]
    n_tb = 200
    tbs_info, coded = [], []
    for i in range(n_tb):
        tbs, qm, G = sizes[i % len(sizes)]
        seg = ol.PortCbsegm()
        assert P.port_cbsegm(C.byref(seg), tbs) == 0 and seg.F == 0
        payload = rng.integers(0, 2, tbs, dtype=np.uint8)
        tb = vec.attach_crc(vec.CRC24A, payload[None, :])[0]
        blocks, pos = [], 0
        for cb in range(seg.C):
            K = seg.K1 if cb < seg.C1 else seg.K2
            rlen = K if seg.C == 1 else K - 24
            blk = tb[pos:pos + rlen]
            pos += rlen
            if seg.C > 1:
                blk = vec.attach_crc(vec.CRC24B, blk[None, :])[0]
            blocks.append(vec.turbo_encode(blk[None, :]))
        tbs_info.append((tbs, qm, G, seg))
        coded.append(blocks)
    dec = P.port_tdec_new()
    sbs = []
    for (tbs, qm, G, seg) in tbs_info:
        sb = ol.PortSoftbuffer()
        P.port_softbuffer_init(C.byref(sb), seg.C)
        sbs.append(sb)
    pool = ctx.harq_pool(n_tb, 13)
    for i in range(n_tb):
        pool.reset(i)
    alive = list(range(n_tb))
    for rv in (0, 2, 3, 1):
        descs, want = [], []
        for i in alive:
            tbs, qm, G, seg = tbs_info[i]
            Gp, gamma = G // qm, (G // qm) % seg.C
            e_parts = []
            for cb in range(seg.C):
                E = qm * (Gp // seg.C) if cb <= seg.C - gamma - 1 else qm * ((Gp + seg.C - 1) // seg.C)
                e_parts.append(vec.rate_match(coded[i][cb], E, rv)[0])
            sigma = (0.55, 0.75, 0.95, 1.2)[i % 4]          # some TBs need one, some several transmissions
            llr = vec.awgn_llr(np.concatenate(e_parts), sigma, 400, rng)
            out = np.zeros(tbs // 8 + 8, np.uint8)
            avg = C.c_float()
            noi = np.zeros(seg.C, np.uint32)
            rc = P.port_decode_tb(dec, C.byref(sbs[i]), tbs, qm, rv, G, llr, out, 10, C.byref(avg), noi)
            want.append((rc, out[: tbs // 8 + 3].copy(), avg.value))
            descs.append(dict(tbs=tbs, qm=qm, rv=rv, e_bits=llr, softbuffer=i))
        res = ctx.decode_tb_batch(pool, descs, 10)
        nxt = []
        for (i, (ret, data, avg), (w_rc, w_out, w_avg)) in zip(alive, res, want):
            tbs, qm, G, seg = tbs_info[i]
            assert ret == w_rc, (rv, i, tbs)
            assert np.array_equal(data[: tbs // 8 + 3], w_out), (rv, i, tbs)
            assert abs(avg - w_avg) < 1e-4, (rv, i, avg, w_avg)
            want_crc = np.array([bool(sbs[i].cb_crc[c]) for c in range(seg.C)])
            assert np.array_equal(pool.cb_crc(i, seg.C).astype(bool), want_crc), (rv, i)
            if ret != 0:
                nxt.append(i)
        if rv == 0:
            assert 0 < len(nxt) < n_tb      # a real mix: some done after the first transmission, some not
        alive = nxt
        if not alive:
            break
    for sb in sbs:
        P.port_softbuffer_free(C.byref(sb))
    P.port_tdec_free(dec)
    pool.close()
