"""Cost of the decoder tiers (development aid): config-3 style batch (K=6144, max 8 half iterations, CRC24B) at two
operating points; run under SRSLTE_B200_SKIP_TIERS / SRSLTE_B200_FORCE_BITS to force tiers.  usage: tier_cost.py [blocks]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
import bench_configs as bc
pkg = ge.load_package(); vec = pkg.vectors
dev = torch.device("cuda", 0); ctx = pkg.Context(0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
K = 6144; n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rng = np.random.default_rng(3)
payload = rng.integers(0, 2, (512, K - 24), dtype=np.uint8)
coded = torch.from_numpy(vec.turbo_encode(vec.attach_crc(vec.CRC24B, payload))).to(dev)
out = torch.zeros((n, K // 8), dtype=torch.uint8, device=dev)
nit = torch.zeros(n, dtype=torch.uint8, device=dev); ok = torch.zeros(n, dtype=torch.uint8, device=dev)
for e_db in (1.5, 4.0):
    llr = bc._noisy(torch, dev, coded, n, vec.harness_sigma(e_db), seed=int(e_db * 10))
    for mode, name in ((pkg.CRC_24B, "crc24b"), (pkg.CRC_NONE, "none")):
        f0 = ctx.fallback_count
        t0 = ctx.tier_counts
        ms = bc._timed(torch, stream, lambda: ctx.tdec_batch_dev(llr.data_ptr(), n, 3 * K + 12, K, 8, out.data_ptr(), K // 8,
                                                                 nit.data_ptr(), ok.data_ptr(), crc_mode=mode, natural=True))
        print(f"e {e_db} crc {name}: {ms:.3f} ms, mean half-its {nit.float().mean().item():.2f}, ms per half-it per 65536 blocks "
              f"{ms / nit.float().mean().item() * 65536 / n:.3f}, exact fallbacks {(ctx.fallback_count - f0) // 5}, "
              f"tiers pure/static/tracked/exact {[ (a - b) // 5 for a, b in zip(ctx.tier_counts, t0)]}")
