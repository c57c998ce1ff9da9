#!/bin/bash
# round-2 GPU call 11 (1 GPU): 3D kernels -- P2G at 6 CTAs/SM, cheaper Newton polar, G2P grid-stride loop with position prefetch (A/B)
mkdir -p gpurun_out
echo start > gpurun_out/r2l_box.txt
MPM_SKIP_HUGE=1 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py tests/test_gpu_fullsize.py tests/test_gpu_deterministic.py -m gpu -q -x -k "3d or 3D or config5 or lift or slab or group or determin or bit" --durations=3 > gpurun_out/r2l_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2l_box.txt
for v in default w0 w8 w2m8; do
  lib=tools/ab/libmpm_$v.so; [ $v = default ] && lib=mpm_flip98a_b200/libmpm.so
  MPM_LIBRARY=$lib timeout 600 python bench.py --workload c5 --steps 40 --warmup 5 --no-cpu --e2e-calls 1 > gpurun_out/r2l_bench_c5_$v.json 2> gpurun_out/r2l_bench_c5_$v.err; echo "bench c5 $v rc=$?" >> gpurun_out/r2l_box.txt
done
timeout 600 python bench.py --workload c2 --steps 40 --warmup 5 --no-cpu --e2e-calls 1 > gpurun_out/r2l_bench_c2.json 2> gpurun_out/r2l_bench_c2.err; echo "bench c2 rc=$?" >> gpurun_out/r2l_box.txt
cat gpurun_out/r2l_box.txt
