for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; echo "N=$n rc=$?"; tail -c 400 gpurun_out/scale_n$n.json
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/scale_ref8.json 2> gpurun_out/scale_ref8.err; echo "ref rc=$?"; tail -c 300 gpurun_out/scale_ref8.json
