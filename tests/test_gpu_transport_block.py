"""Transport-block path on the GPU (rate de-matching -> decode -> CRC -> TB assembly, HARQ) against the
golden vectors produced by the reference's srslte_dlsch_decode2 and against the oracle.  B200 only."""
import ctypes as C

import numpy as np
import pytest

import oracle_libs as ol

pytestmark = pytest.mark.gpu


def _cases(g, prefix):
    return sorted({k.split("_")[0] for k in g if k.startswith(prefix)}, key=lambda s: int(s[1:]))


def test_transport_block_golden_with_harq(ctx, golden):
    """BASELINE config 2 (TBS 75376 = 13 x K=5824, 64QAM) and smaller TBs, rv 0 then rv 2 on the same
    soft buffer: return code, bytes, iteration count and per-block CRC flags equal the reference's."""
    g = golden["tb_vectors"]
    cases = _cases(g, "t")
    pool = ctx.harq_pool(len(cases), 13)
    for sbi, c in enumerate(cases):
        tbs, qm, G, max_it = (int(x) for x in g[f"{c}_par"])
        seg = ol.PortCbsegm()
        ol.port().port_cbsegm(C.byref(seg), tbs)
        pool.reset(sbi)
        for rv in (0, 2):
            (ret, data, avg), = ctx.decode_tb_batch(pool, [dict(tbs=tbs, qm=qm, rv=rv, e_bits=g[f"{c}_rv{rv}_llr"],
                                                                softbuffer=sbi)], max_it)
            want_rc, want_its = (int(x) for x in g[f"{c}_rv{rv}_res"])
            assert ret == want_rc, (c, rv)
            assert round(avg * seg.C) == want_its, (c, rv, avg)
            assert np.array_equal(data[: tbs // 8 + 3], g[f"{c}_rv{rv}_out"]), (c, rv)
            assert np.array_equal(pool.cb_crc(sbi, seg.C), g[f"{c}_rv{rv}_cbcrc"]), (c, rv)
            if ret == 0:
                assert np.array_equal(data[: tbs // 8], g[f"{c}_data"])
    pool.close()


def test_many_transport_blocks_in_one_batch_vs_oracle(ctx, vec):
    """BASELINE config 5 shape: many UL-like TBs of different sizes in ONE call, each against port_decode_tb."""
    P = ol.port()
    rng = np.random.default_rng(8)
    sizes = [(2216, 4, 4800), (6200, 4, 9600), (14112, 4, 28800), (4584, 4, 7200), (1000, 2, 2400), (36696, 6, 60000),
             (3624, 4, 5760), (75376, 6, 90000)]
    tbs_list, want = [], []
    dec = P.port_tdec_new()
    for i, (tbs, qm, G) in enumerate(sizes * 3):
        seg = ol.PortCbsegm()
        assert P.port_cbsegm(C.byref(seg), tbs) == 0 and seg.F == 0
        # build the TB: payload + CRC24A, segmentation with CRC24B, encode + rate match per block (36.212 5.1-5.3)
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
            E = qm * (Gp // seg.C) if cb <= seg.C - gamma - 1 else qm * ((Gp + seg.C - 1) // seg.C)   # encoder's rule
            e_parts.append(vec.rate_match(vec.turbo_encode(blk[None, :]), E, 0)[0])
        e = np.concatenate(e_parts)
        assert e.size == G
        sigma = (0.0, 0.5, 0.8)[i % 3]
        llr = vec.awgn_llr(e, sigma, 100, rng)
        sb = ol.PortSoftbuffer()
        P.port_softbuffer_init(C.byref(sb), seg.C)
        out = np.zeros(tbs // 8 + 8, np.uint8)
        avg = C.c_float()
        noi = np.zeros(seg.C, np.uint32)
        rc = P.port_decode_tb(dec, C.byref(sb), tbs, qm, 0, G, llr, out, 8, C.byref(avg), noi)
        want.append((rc, out[: tbs // 8 + 3].copy(), avg.value, np.packbits(payload)))
        P.port_softbuffer_free(C.byref(sb))
        tbs_list.append(dict(tbs=tbs, qm=qm, rv=0, e_bits=llr, softbuffer=i))
    P.port_tdec_free(dec)
    pool = ctx.harq_pool(len(tbs_list), 13)
    got = ctx.decode_tb_batch(pool, tbs_list, 8)
    n_ok = 0
    for i, ((ret, data, avg), (rc, out, wavg, payload)) in enumerate(zip(got, want)):
        tbs = tbs_list[i]["tbs"]
        assert ret == rc, i
        assert abs(avg - wavg) < 1e-6, (i, avg, wavg)
        assert np.array_equal(data[: tbs // 8 + 3], out), i
        if ret == 0:
            n_ok += 1
            assert np.array_equal(data[: tbs // 8], payload)
    assert n_ok >= len(want) // 2     # the clean and the moderate-noise thirds decode
    pool.close()


def test_tb_argument_errors(ctx):
    pool = ctx.harq_pool(1, 2)
    e = np.zeros(1000, np.int16)
    # more code blocks than the soft buffer holds -> -2; tbs 0 -> 0 (decode_tb returns SUCCESS for empty TBs)
    (ret, _, _), = ctx.decode_tb_batch(pool, [dict(tbs=36696, qm=6, rv=0, e_bits=np.zeros(60000, np.int16), softbuffer=0)], 4)
    assert ret == -2
    (ret, _, _), = ctx.decode_tb_batch(pool, [dict(tbs=0, qm=2, rv=0, e_bits=e, softbuffer=0)], 4)
    assert ret == 0
    (ret, _, _), = ctx.decode_tb_batch(pool, [dict(tbs=1000, qm=2, rv=0, e_bits=e, softbuffer=5)], 4)
    assert ret == -2
    pool.close()


def test_lazy_harq_reset_equals_zero_and_accumulate(ctx, vec):
    """srslte_b200_harq_reset only flags the soft buffer; the next rate de-matching into each of its blocks stores
    instead of adding.  A soft buffer that held a 13-block TB is reset and reused for a 2-block TB, then -- without
    a reset -- receives a second transmission (rv 2): results equal the oracle's zeroed-and-accumulated buffer."""
    P = ol.port()
    rng = np.random.default_rng(31)

    def make(tbs, qm, G, rv, payload=None):
        seg = ol.PortCbsegm()
        assert P.port_cbsegm(C.byref(seg), tbs) == 0 and seg.F == 0
        payload = rng.integers(0, 2, tbs, dtype=np.uint8) if payload is None else payload
        tb = vec.attach_crc(vec.CRC24A, payload[None, :])[0]
        parts, pos = [], 0
        Gp, gamma = G // qm, (G // qm) % seg.C
        for cb in range(seg.C):
            K = seg.K1 if cb < seg.C1 else seg.K2
            rlen = K if seg.C == 1 else K - 24
            blk = tb[pos:pos + rlen]
            pos += rlen
            if seg.C > 1:
                blk = vec.attach_crc(vec.CRC24B, blk[None, :])[0]
            E = qm * (Gp // seg.C) if cb <= seg.C - gamma - 1 else qm * ((Gp + seg.C - 1) // seg.C)
            parts.append(vec.rate_match(vec.turbo_encode(blk[None, :]), E, rv)[0])
        return seg, payload, np.concatenate(parts)

    pool = ctx.harq_pool(1, 13)
    # 1. fill the soft buffer with a big, very noisy TB (fails): leaves garbage in all 13 block buffers
    seg, _, e = make(75376, 6, 90000, 0)
    (ret, _, _), = ctx.decode_tb_batch(pool, [dict(tbs=75376, qm=6, rv=0, e_bits=vec.awgn_llr(e, 1.5, 100, rng), softbuffer=0)], 2)
    assert ret == -1
    # 2. reset, then a 2-block TB in two transmissions at a noise level where one alone fails
    pool.reset(0)
    tbs, qm, G = 8760, 2, 10000
    seg, payload, e0 = make(tbs, qm, G, 0)
    _, _, e2 = make(tbs, qm, G, 2, payload)
    llr0, llr2 = vec.awgn_llr(e0, 1.3, 100, rng), vec.awgn_llr(e2, 1.3, 100, rng)
    dec = P.port_tdec_new()
    sb = ol.PortSoftbuffer()
    P.port_softbuffer_init(C.byref(sb), seg.C)
    for rv, llr in ((0, llr0), (2, llr2)):
        out = np.zeros(tbs // 8 + 8, np.uint8)
        avg = C.c_float()
        noi = np.zeros(seg.C, np.uint32)
        rc = P.port_decode_tb(dec, C.byref(sb), tbs, qm, rv, G, llr, out, 6, C.byref(avg), noi)
        (ret, data, gavg), = ctx.decode_tb_batch(pool, [dict(tbs=tbs, qm=qm, rv=rv, e_bits=llr, softbuffer=0)], 6)
        assert ret == rc, rv
        assert abs(gavg - avg.value) < 1e-6, rv
        assert np.array_equal(data[: tbs // 8 + 3], out[: tbs // 8 + 3]), rv
    P.port_softbuffer_free(C.byref(sb))
    P.port_tdec_free(dec)
    pool.close()


def test_reset_many_and_large_batch_staging(ctx, pkg, vec):
    """srslte_b200_harq_reset_many == one srslte_b200_harq_reset per soft buffer; a batch large enough for the two-half
    staging (helper thread) and the split TB CRC gives the same bytes run after run, and a corrupted TB still fails."""
    import bench_configs as bc
    rng = np.random.default_rng(12)
    sizes = [(2216, 4, 4800), (6200, 4, 9600), (14112, 4, 28800), (4584, 4, 7200), (20616, 4, 36000)]
    descs, pay = [], []
    for i in range(60):
        tbs, qm, G = sizes[i % len(sizes)]
        p, e = bc._make_tb(vec, rng, tbs, qm, G, 0.35, 400)
        pay.append(p)
        descs.append(dict(tbs=tbs, qm=qm, rv=0, e_bits=e, softbuffer=i))
    descs[7]["e_bits"] = (-descs[7]["e_bits"]).astype(np.int16)          # all bits flipped: cannot pass its CRCs
    pool = ctx.harq_pool(60, 13)
    L = pkg.lib()
    first = ctx.decode_tb_batch(pool, descs, 10)
    for i, (ret, data, _) in enumerate(first):
        if i == 7:
            assert ret != 0
        else:
            assert ret == 0 and np.array_equal(np.unpackbits(data[: descs[i]["tbs"] // 8]), pay[i]), i
    # without a reset the good blocks are already decoded; after reset_many everything is decoded again
    assert L.srslte_b200_harq_reset_many(ctx._h, pool._p, None, 60) == 0
    again = ctx.decode_tb_batch(pool, descs, 10)
    for a, b in zip(first, again):
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and a[2] == b[2]
    idx = np.array([3, 9, 11], dtype=np.uint32)
    assert L.srslte_b200_harq_reset_many(ctx._h, pool._p, idx.ctypes.data_as(C.c_void_p), 3) == 0
    assert all(pool.cb_crc(int(i), 1)[0] == 0 for i in idx) and pool.cb_crc(4, 1)[0] == 1
    bad = np.array([60], dtype=np.uint32)
    assert L.srslte_b200_harq_reset_many(ctx._h, pool._p, bad.ctypes.data_as(C.c_void_p), 1) != 0
    pool.close()


def test_pinned_caller_buffers_are_copied_without_staging(ctx, pkg, vec):
    """large TBs whose LLRs sit in pinned host memory take the direct-copy path; the same TBs from pageable memory and a
    mix of both (falls back to staging) give the same bytes"""
    import bench_configs as bc
    rng = np.random.default_rng(31)
    descs, pay, pins = [], [], []
    for i in range(6):
        p, e = bc._make_tb(vec, rng, 75376, 6, 90000, 0.12, 400)
        pay.append(p)
        pa = pkg.PinnedArray(e.shape, np.int16)
        pa.array[:] = e
        pins.append(pa)
        descs.append(dict(tbs=75376, qm=6, rv=0, e_bits=e, softbuffer=i))
    pool = ctx.harq_pool(6, 13)
    L = pkg.lib()

    def run(bufs):
        assert L.srslte_b200_harq_reset_many(ctx._h, pool._p, None, 6) == 0
        arr = (pkg.TbDesc * 6)()
        outs = [np.zeros(75376 // 8 + 8, np.uint8) for _ in range(6)]
        for i in range(6):
            arr[i] = pkg.TbDesc(75376, 6, 0, 90000, i, bufs[i].ctypes.data, outs[i].ctypes.data, 0, 0.0)
        assert L.srslte_b200_decode_tb_batch(ctx._h, pool._p, arr, 6, 10) == 0
        return [(arr[i].ret, outs[i].copy(), arr[i].avg_iterations) for i in range(6)]

    pageable = run([d["e_bits"] for d in descs])
    pinned = run([p.array for p in pins])
    mixed = run([pins[i].array if i % 2 else descs[i]["e_bits"] for i in range(6)])
    for i in range(6):
        assert pageable[i][0] == 0 and np.array_equal(np.unpackbits(pageable[i][1][: 75376 // 8]), pay[i])
        for other in (pinned, mixed):
            assert other[i][0] == 0 and np.array_equal(other[i][1], pageable[i][1]) and other[i][2] == pageable[i][2]
    pool.close()
    for p in pins:
        p.free()
