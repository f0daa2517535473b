"""Config 5 (200 UL transport blocks per subframe) host phase trace: run with SRSLTE_B200_TRACE=1 to see the phases of
srslte_b200_decode_tb_batch on stderr.  usage: SRSLTE_B200_TRACE=1 python tools/cfg5_profile.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import bench_configs as bc
pkg = ge.load_package(); vec = pkg.vectors
rng5 = np.random.default_rng(5)
sizes5 = [(2216, 4, 4800), (6200, 4, 9600), (14112, 4, 28800), (4584, 4, 7200), (3624, 4, 5760), (9144, 4, 14400),
          (1000, 4, 2400), (20616, 4, 36000)]
d5 = []
for i in range(200):
    tbs, qm, G = sizes5[i % len(sizes5)]
    p, e = bc._make_tb(vec, rng5, tbs, qm, G, 0.35, 400)
    d5.append(dict(tbs=tbs, qm=qm, rv=0, e_bits=e))
Lc = pkg.lib(); cx = pkg.Context(0); pl = cx.harq_pool(200, 13)
arr = (pkg.TbDesc * 200)(); keep = []
for i, d in enumerate(d5):
    e = np.ascontiguousarray(d["e_bits"], dtype=np.int16); out = np.zeros(d["tbs"] // 8 + 8, np.uint8); keep += [e, out]
    arr[i] = pkg.TbDesc(d["tbs"], d["qm"], d["rv"], e.shape[0], i, e.ctypes.data, out.ctypes.data, 0, 0.0)
def once():
    Lc.srslte_b200_harq_reset_many(cx._h, pl._p, None, 200)
    return Lc.srslte_b200_decode_tb_batch(cx._h, pl._p, arr, 200, 10)
for _ in range(3): once()
t0 = time.perf_counter()
for _ in range(20): once()
print("per subframe: %.3f ms" % ((time.perf_counter() - t0) / 20 * 1e3))
t0 = time.perf_counter()
for _ in range(20):
    for i in range(200): Lc.srslte_b200_harq_reset(cx._h, pl._p, i)
print("200 resets only: %.3f ms" % ((time.perf_counter() - t0) / 20 * 1e3))
cx.enable_timing(True)
for _ in range(5): once()
cx.synchronize()
for k, name in ((0, 'W16'), (1, 'W8'), (2, 'gen'), (3, 'layout'), (4, 'rm')):
    ms, n = cx.kernel_time(k); print(name, "%.3f ms per call over %d launches" % (ms / 5, n))
