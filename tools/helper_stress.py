"""Stress of the HARQ pool's helper-thread hand-over: config-5 style subframes from two caller threads for N seconds
(each call hands the helper two jobs).  usage: python tools/helper_stress.py [seconds]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import bench_configs as bc
pkg = ge.load_package(); vec = pkg.vectors
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
rng = np.random.default_rng(5)
sizes = [(2216, 4, 4800), (6200, 4, 9600), (14112, 4, 28800), (4584, 4, 7200), (3624, 4, 5760), (9144, 4, 14400), (1000, 4, 2400), (20616, 4, 36000)]
d = []
for i in range(200):
    tbs, qm, G = sizes[i % len(sizes)]
    p, e = bc._make_tb(vec, rng, tbs, qm, G, 0.35, 400)
    d.append(dict(tbs=tbs, qm=qm, rv=0, e_bits=e))
t0 = time.time()
r, bad, _ = bc._tb_rate(pkg, d, 2, secs, 10)
print(f"{r:.0f} subframes/s over {time.time() - t0:.1f} s with 2 caller threads = {2 * r * secs:.0f} helper hand-overs, failed calls {bad}")
