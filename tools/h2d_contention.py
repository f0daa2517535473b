"""Does a running decode kernel slow a concurrent pinned H2D copy down?  (explains the e2e gap to the PCIe rate)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package()
dev = torch.device("cuda", 0)
K, n = 6144, 32768
ctx = pkg.Context(0)
s_comp = torch.cuda.Stream(); s_copy = torch.cuda.Stream()
ctx.set_stream(s_comp.cuda_stream)
llr = (torch.randn((n, 3 * K + 12), device=dev) * 150).to(torch.int16)
out = torch.zeros((n, K // 8), dtype=torch.uint8, device=dev)
nb = 604 << 20
h = torch.empty(nb, dtype=torch.uint8).pin_memory(); d = torch.empty(nb, dtype=torch.uint8, device=dev)
def copy_ms(reps=4, sync=True):
    if sync:
        torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s_copy):
        e0.record(s_copy)
        for _ in range(reps): d.copy_(h, non_blocking=True)
        e1.record(s_copy)
    return e0, e1, reps
def decode(times):
    for _ in range(times):
        ctx.tdec_batch_dev(llr.data_ptr(), n, 3 * K + 12, K, 4, out.data_ptr(), K // 8)
decode(2); torch.cuda.synchronize()
e0, e1, r = copy_ms(); torch.cuda.synchronize()
print(f"H2D alone: {nb * r / e0.elapsed_time(e1) / 1e6:.1f} GB/s")
decode(14)                      # ~60 ms of kernels queued on the compute stream
e0, e1, r = copy_ms(sync=False); torch.cuda.synchronize()
print(f"H2D while the decode kernels run: {nb * r / e0.elapsed_time(e1) / 1e6:.1f} GB/s")
