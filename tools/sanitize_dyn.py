"""compute-sanitizer target for the block-granular early-termination kernel: small launches with refills, cooperative hard
decisions, the aligned (general path) mode and several epochs.  usage: compute-sanitizer --tool memcheck|racecheck python tools/sanitize_dyn.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
pkg = ge.load_package(); vec = pkg.vectors
ctx = pkg.Context(0)
tot = 0
for K, n in ((1024, 700), (6144, 45), (1008, 90), (512, 300), (504, 70)):
    parts = [vec.make_blocks((n + 2) // 3, K, vec.harness_sigma(e), 100.0, seed=K + i)[1] for i, e in enumerate((2.0, 4.5, 7.0))]
    llr = np.ascontiguousarray(np.concatenate(parts)[:n][np.random.default_rng(K).permutation(n)])
    out, nit, ok = ctx.tdec_batch_host(llr, K, 6, crc_mode=pkg.CRC_24B)
    tot += int(ok.sum())
    print(K, n, "ok", int(ok.sum()), "mean half-its %.2f" % nit.mean(), flush=True)
Ks = np.repeat(np.array([2048, 1024, 816, 512], dtype=np.uint32), 21)
llr = np.zeros((len(Ks), 3 * 2048 + 12), np.int16)
for K in (2048, 1024, 816, 512):
    rows = np.nonzero(Ks == K)[0]
    llr[rows, : 3 * K + 12] = vec.make_blocks(len(rows), K, vec.harness_sigma(4.5), 100.0, seed=K)[1]
import ctypes as C
b = pkg.TdecBatch()
b.n_cb = len(Ks); b.long_cb = Ks.ctypes.data_as(C.POINTER(C.c_uint32)); b.uniform_long_cb = 0
b.in_stride = llr.shape[1]; b.out_stride = 256; b.nof_iterations = 6; b.crc_mode = pkg.CRC_24B; b.input_format = 0
out = np.zeros((len(Ks), 256), np.uint8); nit = np.zeros(len(Ks), np.uint8); ok = np.zeros(len(Ks), np.uint8)
assert pkg.lib().srslte_b200_tdec_batch_host(ctx._h, C.byref(b), llr.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                             nit.ctypes.data_as(C.c_void_p), ok.ctypes.data_as(C.c_void_p)) == 0
print("mixed sizes ok", int(ok.sum()), flush=True)
ctx.synchronize()
print("sanitize_dyn done", tot)
