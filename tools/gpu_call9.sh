#!/bin/bash
# round-2 GPU call 9 (1 GPU): 3D G2P kernel (k_g2p3: early stores, on-the-fly re-sort), 3D P2G with plane records; A/B of the G2P occupancy
mkdir -p gpurun_out
echo start > gpurun_out/r2j_box.txt
MPM_SKIP_HUGE=1 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py tests/test_gpu_fullsize.py tests/test_drivers.py -m gpu -q -x -k "3d or 3D or config5 or lift or slab or driver or group" --durations=3 > gpurun_out/r2j_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2j_box.txt
for v in default g6 g8; do
  lib=tools/ab/libmpm_$v.so; [ $v = default ] && lib=mpm_flip98a_b200/libmpm.so
  MPM_LIBRARY=$lib timeout 600 python bench.py --workload c5 --steps 40 --warmup 5 --no-cpu --e2e-calls 1 > gpurun_out/r2j_bench_c5_$v.json 2> gpurun_out/r2j_bench_c5_$v.err; echo "bench c5 $v rc=$?" >> gpurun_out/r2j_box.txt
done
cat gpurun_out/r2j_box.txt
