#!/bin/bash
# round-2 (8 GPUs): the driver's scaling command at N=8 (c4 weak, overlapped, rebalanced) + the 3D scene on 8 slabs
mkdir -p gpurun_out
echo start > gpurun_out/r02n8_box.txt
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 >> gpurun_out/r02n8_box.txt
run() { # name, args...
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 "$@" > gpurun_out/r02_bench_$name.json 2> gpurun_out/r02_bench_$name.err
  rc=$?
  echo "$name rc=$rc" >> gpurun_out/r02n8_box.txt
  return $rc
}
run c4_n8 || run c4_n8_nooverlap --no-overlap
run c5_n8 --workload c5
cat gpurun_out/r02n8_box.txt
