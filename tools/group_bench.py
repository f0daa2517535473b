"""Single-process multi-device entry (srslte_b200_group_*) on all GPUs of the box: results vs one device, the host-to-device
ceiling with every device copying at once, and end-to-end throughput through srslte_b200_group_tdec_batch_host.
usage: python tools/group_bench.py [blocks_per_device]"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); vec = pkg.vectors
K, NIT = 6144, 4
per = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
ndev = torch.cuda.device_count()
bits, llr = vec.make_blocks(1024, K, vec.harness_sigma(1.5), 100.0, seed=7, crc=False)
res = {"devices": ndev, "blocks_per_device": per}
one = pkg.Context(0)
want = one.tdec_batch_host(llr, K, NIT)[0]
one.close()
for nd in sorted({1, 2, 4, ndev} & set(range(1, ndev + 1))):
    g = pkg.Group(nd)
    n = per * nd
    pin = pkg.PinnedArray((n, 3 * K + 12), np.int16)
    out = pkg.PinnedArray((n, K // 8), np.uint8)
    for i in range(0, n, 1024):
        pin.array[i:i + 1024] = llr[: min(1024, n - i)]
    gbs = g.h2d_probe(pin.array.ctypes.data, per * (3 * K + 12) * 2, reps=3)
    for _ in range(2):
        g.tdec_batch_host(pin.array, K, NIT, out=out.array)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        g.tdec_batch_host(pin.array, K, NIT, out=out.array)
    dt = (time.perf_counter() - t0) / reps
    ok = all(np.array_equal(out.array[i:i + 1024][: len(want)], want[: min(1024, n - i)]) for i in range(0, n, 1024 * max(1, n // 8192)))
    e2e_gbs = per * (3 * K + 12) * 2 / dt / 1e9     # per device
    res[f"n{nd}"] = {"h2d_ceiling_gbs_per_device": [round(x, 2) for x in gbs], "e2e_info_gbps": n * K / dt / 1e9,
                     "ms_per_call": dt * 1e3, "h2d_gbs_per_device_in_e2e": round(e2e_gbs, 2),
                     "e2e_frac_of_h2d_ceiling": e2e_gbs / min(gbs), "matches_one_device": bool(ok)}
    if nd > 1:   # shares in proportion to the measured copy rates
        w = g.calibrate()
        for _ in range(2):
            g.tdec_batch_host(pin.array, K, NIT, out=out.array)
        t0 = time.perf_counter()
        for _ in range(reps):
            g.tdec_batch_host(pin.array, K, NIT, out=out.array)
        dtc = (time.perf_counter() - t0) / reps
        okc = all(np.array_equal(out.array[i:i + 1024][: len(want)], want[: min(1024, n - i)]) for i in range(0, n, 1024 * max(1, n // 8192)))
        res[f"n{nd}"]["calibrated"] = {"weights_gbs": [round(x, 2) for x in w], "e2e_info_gbps": n * K / dtc / 1e9,
                                       "ms_per_call": dtc * 1e3, "matches_one_device": bool(okc)}
    pin.free(); out.free(); g.close()
print(json.dumps(res))
