#!/bin/bash
# round-2 GPU call 12 (2 GPUs): slabs over NCCL after the emigrant-by-value fix, measured-cost rebalancing, c5 / strong scaling
mkdir -p gpurun_out
echo start > gpurun_out/r2m_box.txt
timeout 900 python -m pytest tests/test_gpu_deterministic.py tests/test_gpu_slabs.py tests/test_gpu_parity.py -m gpu -q -k "determin or bit or 3d or unsettled" --durations=3 > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2m_box.txt
run() { # name, args...
  name=$1; shift
  timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 "$@" > gpurun_out/r2m_$name.json 2> gpurun_out/r2m_$name.err
  rc=$?
  echo "$name rc=$rc" >> gpurun_out/r2m_box.txt
  return $rc
}
run c4_n2 || run c4_n2_nooverlap --no-overlap
run c5_n2 --workload c5
run c4_n2_strong --scaling strong --warm-substeps 1000
cat gpurun_out/r2m_box.txt
