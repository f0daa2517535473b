#!/bin/bash
# usage (on a multi-GPU box): tools/scale_check.sh <N> [N ...]  -- bench.py under torchrun the way the driver launches it
for n in "$@"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; echo "N=$n rc=$?"; tail -c 300 gpurun_out/scale_n$n.json
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --impl reference --gpus $n --steps 2 --warmup 1 > gpurun_out/scale_ref$n.json 2> gpurun_out/scale_ref$n.err; echo "ref N=$n rc=$?"; tail -c 300 gpurun_out/scale_ref$n.json
done
