"""Kernel time of the stand-alone front end (soft demodulation + descrambling, srslte_b200_demod_descramble_dev) for the
four modulations, 1024 codewords of 15 000 symbols each, and of the kernel fused with rate de-matching (64QAM).
Prints one line per case; not a bench.py contract line."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

pkg = ge.load_package()
dev = torch.device("cuda", 0)
ctx = pkg.Context(0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)


def ktimed(fn, reps=5):
    """device time of the front-end kernels only (the library's own CUDA events, kind 4)"""
    fn(); ctx.synchronize()
    ctx.enable_timing(True)
    for _ in range(reps): fn()
    ctx.synchronize()
    ms, n = ctx.kernel_time(4)
    ctx.enable_timing(False)
    return ms / reps


n_cw, nsym = 1024, 15000
sym = torch.randn((n_cw, nsym, 2), device=dev, dtype=torch.float32) * 0.7
for qm in (() if os.environ.get("FE_BENCH") == "tx" else (2, 4, 6, 8)):
    e = torch.zeros((n_cw, qm * nsym), dtype=torch.int16, device=dev)
    cws = [dict(qm=qm, nof_symbols=nsym, c_init=1 + 7919 * i, sym_offset=i * nsym, llr_offset=i * qm * nsym) for i in range(n_cw)]
    torch.cuda.synchronize()
    ms = ktimed(lambda: ctx.demod_descramble_dev(cws, sym.data_ptr(), e.data_ptr()))
    byt = n_cw * nsym * (8 + 2 * qm)
    print(f"front end Qm={qm}: {n_cw} codewords x {nsym} symbols -> {n_cw * nsym * qm / 1e6:.1f} M LLRs: {ms:.4f} ms kernel time, "
          f"{byt / ms / 1e6:.0f} GB/s of symbol + LLR traffic ({byt / ms / 1e6 / 6536.7 * 100:.0f}% of the measured HBM copy rate), "
          f"{n_cw * nsym * qm / ms / 1e6:.1f} G LLR/s", flush=True)
    if qm == 6:   # PUSCH: the same codewords through the UL-SCH de-interleaver (an index map per LLR)
        cwu = [dict(c, ul_nof_symb=12) for c in cws]
        ms = ktimed(lambda: ctx.demod_descramble_dev(cwu, sym.data_ptr(), e.data_ptr()))
        print(f"front end Qm=6 with the UL-SCH de-interleaver: {ms:.4f} ms kernel time, {byt / ms / 1e6:.0f} GB/s", flush=True)

# ---- TX mirror: turbo encoder + rate matching, 4096 blocks of K = 6144, rv 0, E = 3K + 12 ----------------------------
K, nb = 6144, 4096
N = 3 * K + 12
bits = torch.randint(0, 2, (nb, K), dtype=torch.uint8, device=dev)
eo = torch.zeros((nb, N), dtype=torch.uint8, device=dev)
blocks = [(K, 0, N, i * K, i * N) for i in range(nb)]
ctx.tcod_rm_tx_batch_dev(blocks, bits.data_ptr(), eo.data_ptr()); ctx.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(3): ctx.tcod_rm_tx_batch_dev(blocks, bits.data_ptr(), eo.data_ptr())
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"TX mirror: {nb} blocks K={K} encoded + rate matched: {ms:.3f} ms per call (host item set-up included), "
      f"{nb * K / ms / 1e6:.1f} Gbit/s of payload", flush=True)
