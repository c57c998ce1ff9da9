#!/bin/bash
# round-2 GPU call 11 (1 GPU): fused 3D substep kernel (parity + A/B against the two-kernel path), P2G at 6 CTAs/SM, cheaper Newton polar
mkdir -p gpurun_out
echo start > gpurun_out/r2l_box.txt
MPM_SKIP_HUGE=1 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py tests/test_gpu_fullsize.py tests/test_gpu_deterministic.py -m gpu -q -k "3d or 3D or config5 or lift or slab or group or determin or bit or unsettled" --durations=3 > gpurun_out/r2l_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2l_box.txt
b() { # name lib extra-args
  MPM_LIBRARY=$2 timeout 600 python bench.py --workload c5 --steps 40 --warmup 5 --no-cpu --e2e-calls 1 $3 > gpurun_out/r2l_bench_c5_$1.json 2> gpurun_out/r2l_bench_c5_$1.err; echo "bench c5 $1 rc=$?" >> gpurun_out/r2l_box.txt
}
b fused mpm_flip98a_b200/libmpm.so ""
b fused6 tools/ab/libmpm_s3d6.so ""
b fused4 tools/ab/libmpm_s3d4.so ""
b nofuse mpm_flip98a_b200/libmpm.so "--no-fuse"
b nofuse_w0 tools/ab/libmpm_w0.so "--no-fuse"
timeout 600 python bench.py --workload c2 --steps 40 --warmup 5 --no-cpu --e2e-calls 1 > gpurun_out/r2l_bench_c2.json 2> gpurun_out/r2l_bench_c2.err; echo "bench c2 rc=$?" >> gpurun_out/r2l_box.txt
cat gpurun_out/r2l_box.txt
