#!/bin/bash
# usage (on the GPU box): tools/profile_run.sh <tag>   -> gpurun_out/<tag>_{bench.json,launches.csv,.ncu-rep}
# bench first (never under a profiler), then the ncu launch list of the same command line (short form), then ONE full
# capture of the layout + decode kernels (the first timed step: 3 warm-up steps x 2 kernels are skipped).
tag=$1
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || exit 1
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --e2e-blocks 4096 > gpurun_out/${tag}_plain1.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --e2e-blocks 4096 > gpurun_out/${tag}_ncu1.log 2>&1
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --e2e-blocks 1024 > gpurun_out/${tag}_plain2.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tdec_win_kernel|to_internal_kernel" -s 6 -c 2 -f -o gpurun_out/${tag} \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --e2e-blocks 1024 > gpurun_out/${tag}_ncu2.log 2>&1
ls -la gpurun_out/${tag}*
