"""Per-kernel time of BASELINE config 4 (all 188 block sizes x 64 blocks in one mixed batch)."""
import os, sys, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); vec = pkg.vectors
dev = torch.device("cuda", 0); ctx = pkg.Context(0)
per = int(os.environ.get("CFG4_PER", "64"))
Ks = np.repeat(np.array(vec.ALL_K, dtype=np.uint32), per)
stride = 3 * 6144 + 12
llr = (torch.randn((len(Ks), stride), device=dev) * 120).to(torch.int16)
out = torch.zeros((len(Ks), 768), dtype=torch.uint8, device=dev); nit = torch.zeros(len(Ks), dtype=torch.uint8, device=dev)
b = pkg.TdecBatch(); arrK = np.ascontiguousarray(Ks)
b.n_cb = len(Ks); b.long_cb = arrK.ctypes.data_as(C.POINTER(C.c_uint32)); b.in_stride = stride; b.out_stride = 768
b.nof_iterations = 4; b.crc_mode = pkg.CRC_NONE; b.input_format = 0
L = pkg.lib()
def run():
    assert L.srslte_b200_tdec_batch_dev(ctx._h, C.byref(b), C.c_void_p(llr.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(nit.data_ptr()), C.c_void_p(0)) == 0
run(); run(); ctx.synchronize()
ctx.enable_timing(True)
for _ in range(5): run()
ctx.synchronize()
for k, name in ((0, "W=16 window kernel"), (1, "W=8 window kernel"), (2, "generic kernel (K <= 400)"), (3, "layout")):
    ms, n = ctx.kernel_time(k)
    nb = {0: int((Ks > 800).sum()), 1: int(((Ks > 400) & (Ks <= 800)).sum()), 2: int((Ks <= 400).sum()), 3: len(Ks)}[k]
    bits = {0: int(Ks[Ks > 800].sum()), 1: int(Ks[(Ks > 400) & (Ks <= 800)].sum()), 2: int(Ks[Ks <= 400].sum()), 3: int(Ks.sum())}[k]
    print(f"{name}: {ms / 5:.3f} ms per call, {nb} blocks, {bits / (ms / 5) / 1e6 if ms else 0:.2f} Gbit/s")
