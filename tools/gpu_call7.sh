#!/bin/bash
# round-2 GPU call 7 (2 GPUs): slab protocol v2 over NCCL after the slab_begin fix; c5 / strong scaling
mkdir -p gpurun_out
echo start > gpurun_out/r2h_box.txt
timeout 600 python -m pytest tests/test_gpu_slabs.py -m gpu -q -k "unsettled or group" > gpurun_out/r2h_test_slabs.log 2>&1; echo "slabs test rc=$?" >> gpurun_out/r2h_box.txt
run() { # name, args...
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 "$@" > gpurun_out/r2h_$name.json 2> gpurun_out/r2h_$name.err
  rc=$?
  echo "$name rc=$rc" >> gpurun_out/r2h_box.txt
  return $rc
}
run c4_n2 || run c4_n2_nooverlap --no-overlap
run c5_n2 --workload c5
run c4_n2_strong --scaling strong --warm-substeps 1000
cat gpurun_out/r2h_box.txt
