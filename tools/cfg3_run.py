"""Config-3 style launch for profiling: K=6144, at most 8 half iterations, one operating point.
usage: cfg3_run.py <e_db> <blocks> <crc 0|1> [variant_bits] [reps]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
import bench_configs as bc
pkg = ge.load_package(); vec = pkg.vectors
e_db = float(sys.argv[1]); n = int(sys.argv[2]); mode = pkg.CRC_24B if int(sys.argv[3]) else pkg.CRC_NONE
bits = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = torch.device("cuda", 0); ctx = pkg.Context(0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
ctx.set_variant_bits(bits)
K = 6144
rng = np.random.default_rng(3)
payload = rng.integers(0, 2, (512, K - 24), dtype=np.uint8)
coded = torch.from_numpy(vec.turbo_encode(vec.attach_crc(vec.CRC24B, payload))).to(dev)
out = torch.zeros((n, K // 8), dtype=torch.uint8, device=dev)
nit = torch.zeros(n, dtype=torch.uint8, device=dev); ok = torch.zeros(n, dtype=torch.uint8, device=dev)
llr = bc._noisy(torch, dev, coded, n, vec.harness_sigma(e_db), seed=int(e_db * 10))
ctx.enable_timing(True)
t0 = ctx.tier_counts
ms = bc._timed(torch, stream, lambda: ctx.tdec_batch_dev(llr.data_ptr(), n, 3 * K + 12, K, 8, out.data_ptr(), K // 8,
                                                         nit.data_ptr(), ok.data_ptr(), crc_mode=mode, natural=True), reps=reps)
dec, nl = ctx.kernel_time(0)
print(f"e {e_db} blocks {n} crc {int(sys.argv[3])} bits {bits}: {ms:.3f} ms per call (decode kernel {dec / nl:.3f}), mean half-its "
      f"{nit.float().mean().item():.3f}, ok {ok.float().mean().item():.3f}, tiers {[(a - b) // (reps + 2) for a, b in zip(ctx.tier_counts, t0)]}")
