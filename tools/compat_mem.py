"""Device memory a compat decoder handle (srslte_tdec_init + one block of each regime through srslte_tdec_run_all) takes."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); vec = pkg.vectors
torch.zeros(1, device="cuda"); torch.cuda.synchronize()
h0 = pkg.CompatTdec(6144); h0.run_all(vec.make_blocks(1, 40, 1.0, seed=1)[1][0], 2, 40)   # CUDA context, module load, tables
free0, _ = torch.cuda.mem_get_info()
hs = [pkg.CompatTdec(6144) for _ in range(4)]
for h in hs:
    for K in (6144, 512, 40):
        _b, llr = vec.make_blocks(1, K, 1.0, seed=K)
        h.run_all(llr[0], 4, K)
free1, _ = torch.cuda.mem_get_info()
print(f"4 more srslte_tdec_t handles, each used for K = 6144, 512 and 40: {(free0 - free1) / 4 / 2**20:.1f} MiB of device memory per handle")
