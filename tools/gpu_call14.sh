#!/bin/bash
# round-2 GPU call 14 (2 GPUs): final code over NCCL -- interior launch without migration code, out-of-line emigrants, balance calibrated at the warm state
mkdir -p gpurun_out
echo start > gpurun_out/r2o_box.txt
run() { # name, args...
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 "$@" > gpurun_out/r2o_$name.json 2> gpurun_out/r2o_$name.err
  rc=$?
  echo "$name rc=$rc" >> gpurun_out/r2o_box.txt
  return $rc
}
run c4_n2 || run c4_n2_nooverlap --no-overlap
run c5_n2 --workload c5
run c4_n2_strong --scaling strong --warm-substeps 1000
cat gpurun_out/r2o_box.txt
