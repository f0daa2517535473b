"""Device-resident timings of the BASELINE.json configurations that are not the bench.py headline:
config 3 (65536 x K=6144, max 8 half iterations, CRC24B early termination, two operating points) and
config 4 (all 188 LTE block sizes in one mixed batch).  Prints one line per case; not a bench.py contract line."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

pkg = ge.load_package(); vec = pkg.vectors
dev = torch.device("cuda", 0)
ctx = pkg.Context(0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)


def noisy(coded, n, sigma, seed, scale=100.0):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    L = coded.shape[1]
    out = torch.empty((n, L), dtype=torch.int16, device=dev)
    for i in range(0, n, 4096):
        m = min(4096, n - i)
        idx = torch.arange(i, i + m, device=dev) % coded.shape[0]
        rx = coded[idx].to(torch.float32) * 2 - 1 + sigma * torch.randn((m, L), device=dev, generator=g)
        out[i:i + m] = torch.trunc(scale * rx).clamp_(-32768, 32767).to(torch.int16)
    return out


def timed(fn, reps=3):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ---- config 3 ---------------------------------------------------------------------------------------
K, n = 6144, int(os.environ.get("CFG3_BLOCKS", "65536"))
rng = np.random.default_rng(3)
payload = rng.integers(0, 2, (512, K - 24), dtype=np.uint8)
bits = vec.attach_crc(vec.CRC24B, payload)
coded = torch.from_numpy(vec.turbo_encode(bits)).to(dev)
out = torch.zeros((n, K // 8), dtype=torch.uint8, device=dev)
nit = torch.zeros(n, dtype=torch.uint8, device=dev); ok = torch.zeros(n, dtype=torch.uint8, device=dev)
for e_db in (1.5, 4.0):
    llr = noisy(coded, n, vec.harness_sigma(e_db), seed=int(e_db * 10))
    ms = timed(lambda: ctx.tdec_batch_dev(llr.data_ptr(), n, 3 * K + 12, K, 8, out.data_ptr(), K // 8, nit.data_ptr(),
                                          ok.data_ptr(), crc_mode=pkg.CRC_24B, natural=True))
    good = (out[ok == 1][:, : (K - 24) // 8].cpu().numpy() ==
            np.packbits(payload, axis=1)[(torch.nonzero(ok == 1).flatten() % 512).cpu().numpy()]).all()
    print(f"config3 harness -e {e_db}: {n} blocks K=6144 max 8 half-its CRC24B: {ms:.2f} ms/step, "
          f"{n * (K - 24) / ms / 1e6:.2f} Gbit/s payload, mean half-its {nit.float().mean().item():.2f}, "
          f"crc ok {ok.float().mean().item() * 100:.1f}%, decoded payloads correct: {bool(good)}, "
          f"exact fallbacks so far {ctx.fallback_count}")
    del llr

# ---- config 4 ---------------------------------------------------------------------------------------
per = 64
Ks = np.repeat(np.array(vec.ALL_K, dtype=np.uint32), per)
stride = 3 * 6144 + 12
llr = torch.zeros((len(Ks), stride), dtype=torch.int16, device=dev)
for K4 in vec.ALL_K:
    b4 = rng.integers(0, 2, (8, K4), dtype=np.uint8)
    c4 = torch.from_numpy(vec.turbo_encode(b4)).to(dev)
    rows = np.nonzero(Ks == K4)[0]
    llr[rows[0]:rows[-1] + 1, : 3 * K4 + 12] = noisy(c4, per, vec.harness_sigma(4.0), seed=K4)
out4 = np.zeros((len(Ks), 768), np.uint8)
t0 = time.perf_counter()
L = pkg.lib()
import ctypes as C
b = pkg.TdecBatch()
# host entry with per-block K (the python wrapper takes one K): go through the raw C ABI
arrK = np.ascontiguousarray(Ks)
b.n_cb = len(Ks); b.long_cb = arrK.ctypes.data_as(C.POINTER(C.c_uint32)); b.uniform_long_cb = 0
b.in_stride = stride; b.out_stride = 768; b.nof_iterations = 4; b.crc_mode = pkg.CRC_NONE; b.input_format = 0
outd = torch.zeros((len(Ks), 768), dtype=torch.uint8, device=dev)
nitd = torch.zeros(len(Ks), dtype=torch.uint8, device=dev)
def run4():
    rc = L.srslte_b200_tdec_batch_dev(ctx._h, C.byref(b), C.c_void_p(llr.data_ptr()), C.c_void_p(outd.data_ptr()),
                                      C.c_void_p(nitd.data_ptr()), C.c_void_p(0))
    assert rc == 0, rc
ms = timed(run4)
print(f"config4 mixed batch: {len(Ks)} blocks = 188 sizes x {per}, 4 half-its: {ms:.3f} ms/step, "
      f"{float(Ks.sum()) / ms / 1e6:.2f} Gbit/s")

# ---- front end: soft demodulation + descrambling, alone and fused into rate de-matching ----------
n_cw, nsym, qm = 1024, 15000, 6        # config 2's codeword: 100 PRB, 64QAM, 90000 bits, 13 blocks of K = 5824
sym = torch.randn((n_cw, nsym, 2), device=dev, dtype=torch.float32) * 0.7
e = torch.zeros((n_cw, qm * nsym), dtype=torch.int16, device=dev)
cws = [dict(qm=qm, nof_symbols=nsym, c_init=1 + 7919 * i, sym_offset=i * nsym, llr_offset=i * qm * nsym) for i in range(n_cw)]
def ktimed(fn, reps=3):
    """device time of the front-end kernels only (the library's own CUDA events, kind 4)"""
    fn(); ctx.synchronize()
    ctx.enable_timing(True)
    for _ in range(reps): fn()
    ctx.synchronize()
    ms, n = ctx.kernel_time(4)
    ctx.enable_timing(False)
    return ms / reps


ms = ktimed(lambda: ctx.demod_descramble_dev(cws, sym.data_ptr(), e.data_ptr()))
byt = n_cw * nsym * (8 + 2 * qm)
print(f"front end: {n_cw} codewords x {nsym} symbols 64QAM -> {n_cw * nsym * qm / 1e6:.1f} M LLRs: {ms:.3f} ms kernel time, "
      f"{byt / ms / 1e6:.0f} GB/s of symbol + LLR traffic ({byt / ms / 1e6 / 6536.7 * 100:.0f}% of the measured HBM copy rate), "
      f"{n_cw * nsym * qm / ms / 1e6:.1f} G LLR/s")
wl = 18624
work = torch.zeros((n_cw * 13, wl), dtype=torch.int16, device=dev)
blocks = []
for i in range(n_cw):
    rp = 0
    for cb in range(13):
        E = 6918 if cb <= 2 else 6924     # sch.c:324-334 for G = 90000, C = 13
        blocks.append((5824, 0, i, rp, E, (i * 13 + cb) * wl))
        rp += E
ms = ktimed(lambda: ctx.demod_rm_rx_batch_dev(cws, blocks, sym.data_ptr(), work.data_ptr()))
print(f"front end fused with rate de-matching: {len(blocks)} blocks K=5824: {ms:.3f} ms kernel time, "
      f"{n_cw * 90000 / ms / 1e6:.1f} G LLR/s")
e2 = e.clone()
rmb = [(5824, 0, b[2] * 90000 + b[3], b[4], b[5]) for b in blocks]
ms2 = ktimed(lambda: (ctx.demod_descramble_dev(cws, sym.data_ptr(), e2.data_ptr()),
                      ctx.rm_rx_batch_dev(rmb, e2.data_ptr(), work.data_ptr())))
print(f"front end unfused (demod kernel + rm kernel through the e array): {ms2:.3f} ms kernel time")

# ---- config 5: 200 UL transport blocks per 1 ms subframe (16QAM, mixed sizes), host LLRs in, TB bytes out ----------
import ctypes as C
rng5 = np.random.default_rng(5)
sizes5 = [(2216, 4, 4800), (6200, 4, 9600), (14112, 4, 28800), (4584, 4, 7200), (3624, 4, 5760), (9144, 4, 14400),
          (1000, 4, 2400), (20616, 4, 36000)]
tbs5 = []
for i in range(200):
    tbs, qm, G = sizes5[i % len(sizes5)]
    payload = rng5.integers(0, 2, tbs, dtype=np.uint8)
    tb = vec.attach_crc(vec.CRC24A, payload[None, :])[0]
    # segmentation (36.212 5.1.2) through the library's own table helper
    from math import ceil
    B = tbs + 24
    Cn = 1 if B <= 6144 else ceil(B / (6144 - 24))
    Bp = B if Cn == 1 else B + Cn * 24
    Kp = next(k for k in vec.ALL_K if Cn * k >= Bp)
    assert Cn * Kp == Bp, "sizes chosen without filler bits"
    e_parts, pos = [], 0
    Gp, gamma = G // qm, (G // qm) % Cn
    for cb in range(Cn):
        rlen = Kp if Cn == 1 else Kp - 24
        blk = tb[pos:pos + rlen]; pos += rlen
        if Cn > 1:
            blk = vec.attach_crc(vec.CRC24B, blk[None, :])[0]
        E = qm * (Gp // Cn) if cb <= Cn - gamma - 1 else qm * ((Gp + Cn - 1) // Cn)
        e_parts.append(vec.rate_match(vec.turbo_encode(blk[None, :]), E, 0)[0])
    e5 = np.concatenate(e_parts)
    tbs5.append(dict(tbs=tbs, qm=qm, rv=0, e_bits=vec.awgn_llr(e5, 0.35, 400, rng5), softbuffer=i))
import threading


def subframe_rate(n_threads, seconds=1.5):
    """Sustained subframes/s through the raw C ABI: every thread owns a context, a HARQ pool and its descriptors and
    per subframe resets its 200 HARQ buffers and decodes its 200 TBs (ctypes releases the GIL during the calls)."""
    Lc = pkg.lib()
    workers = []
    for _ in range(n_threads):
        cx = pkg.Context(0)
        pl = cx.harq_pool(200, 13)
        arr = (pkg.TbDesc * 200)()
        keep = []
        for i, d in enumerate(tbs5):
            e = np.ascontiguousarray(d["e_bits"], dtype=np.int16)
            out = np.zeros(d["tbs"] // 8 + 8, np.uint8)
            keep += [e, out]
            arr[i] = pkg.TbDesc(d["tbs"], d["qm"], d["rv"], e.shape[0], d["softbuffer"], e.ctypes.data, out.ctypes.data, 0, 0.0)
        workers.append((cx, pl, arr, keep))
    counts = [0] * n_threads
    stop = [False]

    def loop(w):
        cx, pl, arr, _ = workers[w]
        while not stop[0]:
            for i in range(200):
                Lc.srslte_b200_harq_reset(cx._h, pl._p, i)
            rc = Lc.srslte_b200_decode_tb_batch(cx._h, pl._p, arr, 200, 10)
            assert rc == 0 and all(arr[i].ret == 0 for i in range(0, 200, 37))
            counts[w] += 1
    for w in range(n_threads):          # warm-up
        cx, pl, arr, _ = workers[w]
        for i in range(200): Lc.srslte_b200_harq_reset(cx._h, pl._p, i)
        Lc.srslte_b200_decode_tb_batch(cx._h, pl._p, arr, 200, 10)
    ths = [threading.Thread(target=loop, args=(w,)) for w in range(n_threads)]
    t0 = time.perf_counter()
    for th in ths: th.start()
    time.sleep(seconds)
    stop[0] = True
    for th in ths: th.join()
    dt = time.perf_counter() - t0
    for cx, pl, _, _ in workers:
        pl.close()
    return sum(counts) / dt


bits5 = sum(d["tbs"] for d in tbs5)
for nt in (1, 2, 3):
    r = subframe_rate(nt)
    print(f"config5: 200 TBs (16QAM, {bits5} payload bits, {sum(len(d['e_bits']) for d in tbs5)} LLRs) per subframe, host LLRs in, "
          f"TB bytes out, 200 HARQ resets per subframe, {nt} caller thread(s) on one GPU: {r:.0f} subframes/s "
          f"({bits5 * r / 1e9:.2f} Gbit/s payload)")
