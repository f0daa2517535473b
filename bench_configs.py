"""The BASELINE.json configurations next to the headline one, measured inside bench.py (key "configs" of the JSON line).

  config 1  one K=6144 block through the unchanged srslte_tdec_run_all name (latency)
  config 2  pdsch_test's transport block: 100 PRB, 64QAM, MCS 28 -> 13 code blocks of K=5824: rate de-matching + turbo
            decode + CRC24B/24A through srslte_b200_decode_tb_batch, many TBs per call, host LLRs in, TB bytes out
  config 3  65 536 x K=6144, at most 8 half iterations, CRC24B early termination, at harness -e 1.5 and -e 4.0
  config 4  all 188 LTE block sizes x 64 blocks in one mixed batch, 4 half iterations
  config 5  200 UL transport blocks (16QAM, mixed sizes) per 1 ms subframe, sustained subframes per second
  frontend  soft demodulation + descrambling alone, through the UL-SCH de-interleaver and fused with rate
            de-matching: kernel times

Every case also decodes a small sample with the oracle (tests/oracle_libs.py: the scalar port; a CHECKER here, never
timed) and compares bit for bit: "parity" in each entry.
"""
import ctypes as C
import os
import sys
import threading
import time
from math import ceil

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def _noisy(torch, dev, coded, n, sigma, seed, scale=100.0):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    L = coded.shape[1]
    out = torch.empty((n, L), dtype=torch.int16, device=dev)
    for i in range(0, n, 4096):
        m = min(4096, n - i)
        idx = torch.arange(i, i + m, device=dev) % coded.shape[0]
        rx = coded[idx].to(torch.float32) * 2 - 1 + sigma * torch.randn((m, L), device=dev, generator=g)
        out[i:i + m] = torch.trunc(scale * rx).clamp_(-32768, 32767).to(torch.int16)
    return out


def _timed(torch, stream, fn, reps=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _crc_mode_expect(ol, llr1, K, max_it):
    """what srslte's sch.c loop gives for one block: (bytes, half iterations, crc ok), from the oracle's trace"""
    P = ol.port()
    by, _, _ = ol.port_trace(llr1, K, max_it)
    for it in range(max_it):
        if P.port_crc_bytes(ol.CRC24B, by[it].copy(), K) == 0:
            return by[it], it + 1, 1
    return by[max_it - 1], max_it, 0


def _segment(vec, tbs):
    B = tbs + 24
    Cn = 1 if B <= 6144 else ceil(B / (6144 - 24))
    Bp = B if Cn == 1 else B + Cn * 24
    Kp = next(k for k in vec.ALL_K if Cn * k >= Bp)
    assert Cn * Kp == Bp, "sizes chosen without filler bits"
    return Cn, Kp


def _make_tb(vec, rng, tbs, qm, G, sigma, scale):
    """one transport block: payload -> CRC24A -> segmentation -> per-block CRC24B -> turbo code -> rate matching -> LLRs"""
    payload = rng.integers(0, 2, tbs, dtype=np.uint8)
    tb = vec.attach_crc(vec.CRC24A, payload[None, :])[0]
    Cn, Kp = _segment(vec, tbs)
    e_parts, pos = [], 0
    Gp, gamma = G // qm, (G // qm) % Cn
    for cb in range(Cn):
        rlen = Kp if Cn == 1 else Kp - 24
        blk = tb[pos:pos + rlen]
        pos += rlen
        if Cn > 1:
            blk = vec.attach_crc(vec.CRC24B, blk[None, :])[0]
        E = qm * (Gp // Cn) if cb <= Cn - gamma - 1 else qm * ((Gp + Cn - 1) // Cn)
        e_parts.append(vec.rate_match(vec.turbo_encode(blk[None, :]), E, 0)[0])
    return payload, vec.awgn_llr(np.concatenate(e_parts), sigma, scale, rng)


def _tb_rate(pkg, descs, n_threads, seconds, max_it, resets=True, pinned=False):
    """sustained calls per second of srslte_b200_decode_tb_batch over `descs` through the raw C ABI, one context +
    HARQ pool per caller thread (ctypes releases the GIL during the calls)"""
    Lc = pkg.lib()
    n = len(descs)
    workers = []
    for _ in range(n_threads):
        cx = pkg.Context(0)
        pl = cx.harq_pool(n, 13)
        arr = (pkg.TbDesc * n)()
        keep = []
        for i, d in enumerate(descs):
            e = np.ascontiguousarray(d["e_bits"], dtype=np.int16)
            if pinned:   # the caller's LLR buffers in pinned host memory: large TBs are then copied without staging
                pa = pkg.PinnedArray(e.shape, np.int16)
                pa.array[:] = e
                keep.append(pa)
                e = pa.array
            out = np.zeros(d["tbs"] // 8 + 8, np.uint8)
            keep += [e, out]
            arr[i] = pkg.TbDesc(d["tbs"], d["qm"], d["rv"], e.shape[0], i, e.ctypes.data, out.ctypes.data, 0, 0.0)
        workers.append((cx, pl, arr, keep))
    counts = [0] * n_threads
    stop = [False]
    bad = [0]

    def once(w):
        cx, pl, arr, _ = workers[w]
        if resets:   # every TB is a new transmission: reset all soft buffers (one C call, like a MAC looping in C)
            Lc.srslte_b200_harq_reset_many(cx._h, pl._p, None, n)
        rc = Lc.srslte_b200_decode_tb_batch(cx._h, pl._p, arr, n, max_it)
        if rc != 0 or any(arr[i].ret != 0 for i in range(0, n, max(1, n // 7))):
            bad[0] += 1

    def loop(w):
        while not stop[0]:
            once(w)
            counts[w] += 1
    for w in range(n_threads):
        once(w)
    ths = [threading.Thread(target=loop, args=(w,)) for w in range(n_threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    time.sleep(seconds)
    stop[0] = True
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    outs = [np.array(workers[0][3][(3 if pinned else 2) * i + (2 if pinned else 1)]) for i in range(n)]
    for cx, pl, _, _ in workers:
        pl.close()
        cx.close()
    return sum(counts) / dt, bad[0], outs


def run_configs(pkg, ctx, torch, dev, stream, quick=False):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_libs as ol  # the checker
    vec = pkg.vectors
    res = {}
    K = 6144
    rng = np.random.default_rng(3)

    # ---- config 1: one block through the reference's own entry point name ----
    L = pkg.lib()
    bits1, llr1 = vec.make_blocks(4, K, vec.harness_sigma(1.5), 100.0, seed=11, crc=False)
    want1 = ol.port_run_all(llr1, K, 4)
    h = pkg.CompatTdec(K) if hasattr(pkg, "CompatTdec") else None
    if h is not None:
        outs = [h.run_all(llr1[i], 4, K) for i in range(4)]
        t0 = time.perf_counter()
        reps = 10 if quick else 40
        for r in range(reps):
            h.run_all(llr1[r % 4], 4, K)
        us = (time.perf_counter() - t0) / reps * 1e6
        h.close()
        res["config1_single_block_srslte_tdec_run_all"] = {
            "K": K, "nof_iterations": 4, "latency_us_per_call": us, "info_mbps": K / us,
            "parity": bool(all(np.array_equal(outs[i], want1[i]) for i in range(4))), "parity_sample": "4 blocks vs the oracle port"}

    # ---- config 3 ----
    n3 = 8192 if quick else 65536
    payload = rng.integers(0, 2, (512, K - 24), dtype=np.uint8)
    bits = vec.attach_crc(vec.CRC24B, payload)
    coded = torch.from_numpy(vec.turbo_encode(bits)).to(dev)
    out = torch.zeros((n3, K // 8), dtype=torch.uint8, device=dev)
    nit = torch.zeros(n3, dtype=torch.uint8, device=dev)
    ok = torch.zeros(n3, dtype=torch.uint8, device=dev)
    c3 = {}
    for e_db in (1.5, 4.0):
        llr = _noisy(torch, dev, coded, n3, vec.harness_sigma(e_db), seed=int(e_db * 10))
        ms = _timed(torch, stream, lambda: ctx.tdec_batch_dev(llr.data_ptr(), n3, 3 * K + 12, K, 8, out.data_ptr(), K // 8,
                                                              nit.data_ptr(), ok.data_ptr(), crc_mode=pkg.CRC_24B, natural=True))
        sample = llr[:6].cpu().numpy()
        o_s, n_s, k_s = out[:6].cpu().numpy(), nit[:6].cpu().numpy(), ok[:6].cpu().numpy()
        par = True
        for i in range(6):
            wb, wn, wk = _crc_mode_expect(ol, sample[i], K, 8)
            par = par and np.array_equal(o_s[i], wb) and int(n_s[i]) == wn and int(k_s[i]) == wk
        mean_it = float(nit.float().mean().item())
        c3[f"harness_e{e_db}"] = {"ms_per_step": ms, "payload_gbps": n3 * (K - 24) / ms / 1e6, "mean_half_iterations": mean_it,
                                  "crc_ok_fraction": float(ok.float().mean().item()), "parity": bool(par),
                                  "parity_sample": "6 blocks: bytes, half iterations and CRC flag vs the oracle port"}
        del llr
    c3["blocks"] = n3
    c3["time_ratio_e4.0_over_e1.5"] = c3["harness_e4.0"]["ms_per_step"] / c3["harness_e1.5"]["ms_per_step"]
    c3["half_iteration_ratio_e4.0_over_e1.5"] = c3["harness_e4.0"]["mean_half_iterations"] / c3["harness_e1.5"]["mean_half_iterations"]
    res["config3_batched_k6144_8it_crc_early_termination"] = c3
    del out, nit, ok, coded

    # ---- config 4 ----
    per = 16 if quick else 64
    Ks = np.repeat(np.array(vec.ALL_K, dtype=np.uint32), per)
    stride = 3 * 6144 + 12
    llr = torch.zeros((len(Ks), stride), dtype=torch.int16, device=dev)
    for K4 in vec.ALL_K:
        b4 = rng.integers(0, 2, (8, K4), dtype=np.uint8)
        c4 = torch.from_numpy(vec.turbo_encode(b4)).to(dev)
        rows = np.nonzero(Ks == K4)[0]
        llr[rows[0]:rows[-1] + 1, : 3 * K4 + 12] = _noisy(torch, dev, c4, per, vec.harness_sigma(4.0), seed=K4)
    b = pkg.TdecBatch()
    arrK = np.ascontiguousarray(Ks)
    b.n_cb = len(Ks); b.long_cb = arrK.ctypes.data_as(C.POINTER(C.c_uint32)); b.uniform_long_cb = 0
    b.in_stride = stride; b.out_stride = 768; b.nof_iterations = 4; b.crc_mode = pkg.CRC_NONE; b.input_format = 0
    outd = torch.zeros((len(Ks), 768), dtype=torch.uint8, device=dev)
    nitd = torch.zeros(len(Ks), dtype=torch.uint8, device=dev)

    def run4():
        rc = L.srslte_b200_tdec_batch_dev(ctx._h, C.byref(b), C.c_void_p(llr.data_ptr()), C.c_void_p(outd.data_ptr()),
                                          C.c_void_p(nitd.data_ptr()), C.c_void_p(0))
        assert rc == 0, rc
    ms = _timed(torch, stream, run4)
    par, checked = True, 0
    for K4 in vec.ALL_K[::9]:      # every ninth size: 21 sizes, first block of each
        row = int(np.nonzero(Ks == K4)[0][0])
        want = ol.port_run_all(llr[row:row + 1, : 3 * K4 + 12].cpu().numpy(), K4, 4)[0]
        par = par and np.array_equal(outd[row, : K4 // 8].cpu().numpy(), want)
        checked += 1
    res["config4_mixed_188_sizes"] = {"blocks": int(len(Ks)), "per_size": per, "nof_iterations": 4, "ms_per_step": ms,
                                      "info_gbps": float(Ks.sum()) / ms / 1e6, "parity": bool(par),
                                      "parity_sample": f"{checked} block sizes vs the oracle port"}
    del llr, outd, nitd

    # ---- config 2: pdsch_test's TB, many per call ----
    rng2 = np.random.default_rng(2)
    n2 = 16 if quick else 64
    d2 = []
    pay2 = []
    for i in range(n2):
        p, e = _make_tb(vec, rng2, 75376, 6, 90000 - 90000 % 6, 0.12, 400)   # 13 x K=5824, G = 15000 REs x 6
        pay2.append(p)
        d2.append(dict(tbs=75376, qm=6, rv=0, e_bits=e))
    r2p, bad2p, outs2p = _tb_rate(pkg, d2, 1, 0.8 if quick else 1.5, 10)
    r2, bad2, outs2 = _tb_rate(pkg, d2, 1, 0.8 if quick else 1.5, 10, pinned=True)
    par2 = all(np.array_equal(np.unpackbits(outs2[i][:75376 // 8]), pay2[i]) and
               np.array_equal(np.unpackbits(outs2p[i][:75376 // 8]), pay2[i]) for i in range(n2))
    bad2 += bad2p
    res["config2_pdsch_tb_13x5824"] = {"tbs": 75376, "code_blocks": 13, "K": 5824, "tbs_per_call": n2, "calls_per_s": r2,
                                       "input": "the callers' LLR buffers in pinned host memory (copied without staging)",
                                       "calls_per_s_pageable_input": r2p,
                                       "tb_per_s": r2 * n2, "payload_gbps": r2 * n2 * 75376 / 1e9, "failed_calls": bad2,
                                       "parity": bool(par2 and bad2 == 0),
                                       "parity_sample": "every TB's bytes vs the transmitted payload (CRC24A checked by the library)",
                                       "path": "host LLRs -> H2D -> rate de-matching -> decode (CRC24B early termination) -> TB bytes"}

    # ---- config 5 ----
    rng5 = np.random.default_rng(5)
    sizes5 = [(2216, 4, 4800), (6200, 4, 9600), (14112, 4, 28800), (4584, 4, 7200), (3624, 4, 5760), (9144, 4, 14400),
              (1000, 4, 2400), (20616, 4, 36000)]
    d5, pay5 = [], []
    for i in range(200):
        tbs, qm, G = sizes5[i % len(sizes5)]
        p, e = _make_tb(vec, rng5, tbs, qm, G, 0.35, 400)
        pay5.append(p)
        d5.append(dict(tbs=tbs, qm=qm, rv=0, e_bits=e))
    bits5 = sum(d["tbs"] for d in d5)
    c5 = {"tbs_per_subframe": 200, "payload_bits_per_subframe": bits5, "llrs_per_subframe": int(sum(len(d["e_bits"]) for d in d5))}
    par5 = True
    for nt in ((1,) if quick else (1, 2)):
        r5, bad5, outs5 = _tb_rate(pkg, d5, nt, 0.8 if quick else 1.5, 10)
        par5 = par5 and bad5 == 0 and all(np.array_equal(np.unpackbits(outs5[i][:d5[i]["tbs"] // 8]), pay5[i]) for i in range(200))
        c5[f"subframes_per_s_{nt}_caller_thread"] = r5
    c5["host_threads"] = "each caller thread owns a context + HARQ pool; a pool has one helper thread (half of the staging copy and of the TB CRC24A)"
    c5["parity"] = bool(par5)
    c5["parity_sample"] = "every TB's bytes vs the transmitted payload"
    res["config5_200_ul_tbs_per_subframe"] = c5

    # ---- the stage in front of the path: kernel times ----
    try:
        res["frontend_64qam"] = _frontend(pkg, ctx, torch, dev, ol, quick)
    except Exception as e:  # never lose the line over an auxiliary entry
        res["frontend_64qam"] = {"error": f"{type(e).__name__}: {e}"}
    return res


def _frontend(pkg, ctx, torch, dev, ol, quick):
    """Soft demodulation + descrambling of 1024 codewords x 15 000 64QAM symbols (config 2's codeword), alone, through the
    UL-SCH de-interleaver, and fused with the rate de-matching of their 13 312 code blocks; device time of the kernels
    from the library's own CUDA events.  Parity: two codewords against the oracle port."""
    n_cw, nsym, qm = (128 if quick else 1024), 15000, 6
    sym = torch.randn((n_cw, nsym, 2), device=dev, dtype=torch.float32) * 0.7
    e = torch.zeros((n_cw, qm * nsym), dtype=torch.int16, device=dev)
    cws = [dict(qm=qm, nof_symbols=nsym, c_init=1 + 7919 * i, sym_offset=i * nsym, llr_offset=i * qm * nsym) for i in range(n_cw)]

    def ktimed(fn, reps=5):
        fn(); ctx.synchronize()
        ctx.enable_timing(True)
        for _ in range(reps):
            fn()
        ctx.synchronize()
        ms, _n = ctx.kernel_time(4)
        ctx.enable_timing(False)
        return ms / reps

    torch.cuda.synchronize()
    ms_plain = ktimed(lambda: ctx.demod_descramble_dev(cws, sym.data_ptr(), e.data_ptr()))
    got = e[:2].cpu().numpy()
    sym_h = sym[:2].cpu().numpy()
    ok = all(np.array_equal(got[i], ol.port_demod_descramble(qm, sym_h[i].copy().view(np.complex64).reshape(-1), cws[i]["c_init"]))
             for i in range(2))
    cwu = [dict(c, ul_nof_symb=12) for c in cws]
    ms_ul = ktimed(lambda: ctx.demod_descramble_dev(cwu, sym.data_ptr(), e.data_ptr()))
    got = e[:2].cpu().numpy()
    ok_ul = all(np.array_equal(got[i], ol.port_ulsch_deinterleave(
        ol.port_demod_descramble(qm, sym_h[i].copy().view(np.complex64).reshape(-1), cws[i]["c_init"]), qm, 12)) for i in range(2))
    wl = 18624
    work = torch.zeros((n_cw * 13, wl), dtype=torch.int16, device=dev)
    blocks = []
    for i in range(n_cw):
        rp = 0
        for cb in range(13):
            E = 6918 if cb <= 2 else 6924     # sch.c:324-334 for G = 90000, C = 13
            blocks.append((5824, 0, i, rp, E, (i * 13 + cb) * wl))
            rp += E
    ms_fused = ktimed(lambda: ctx.demod_rm_rx_batch_dev(cws, blocks, sym.data_ptr(), work.data_ptr()))
    byt = n_cw * nsym * (8 + 2 * qm)
    return {"codewords": n_cw, "symbols_per_codeword": nsym, "qm": qm, "llrs": n_cw * nsym * qm,
            "demod_descramble_ms": ms_plain, "demod_descramble_gbs": byt / ms_plain / 1e6,
            "demod_descramble_frac_of_hbm_copy_rate": byt / ms_plain / 1e6 / 6536.7,
            "demod_descramble_pusch_deinterleaver_ms": ms_ul,
            "fused_with_rate_dematching_ms": ms_fused, "fused_code_blocks": len(blocks),
            "timing": "library CUDA events around the kernels, 5 back-to-back launches",
            "parity": bool(ok and ok_ul), "parity_sample": "2 codewords, plain and through the UL-SCH de-interleaver, vs the oracle port"}
