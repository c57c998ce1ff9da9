#!/bin/bash
# round-2 GPU call 13 (1 GPU): 3D G2P with the shared-memory node tile (CTA per chunk, software pipeline) A/B, 3D P2G prefetch A/B
mkdir -p gpurun_out
echo start > gpurun_out/r2n_box.txt
MPM_SKIP_HUGE=1 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py tests/test_gpu_fullsize.py tests/test_gpu_deterministic.py -m gpu -q -k "3d or 3D or config5 or lift or slab or group or determin or bit or unsettled" --durations=3 > gpurun_out/r2n_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2n_box.txt
b() { # name lib extra-args
  MPM_LIBRARY=$2 timeout 600 python bench.py --workload c5 --steps 40 --warmup 5 --no-cpu --e2e-calls 1 $3 > gpurun_out/r2n_bench_c5_$1.json 2> gpurun_out/r2n_bench_c5_$1.err; echo "bench c5 $1 rc=$?" >> gpurun_out/r2n_box.txt
}
b tile7 mpm_flip98a_b200/libmpm.so ""
b tile6 tools/ab/libmpm_t6.so ""
b tile5 tools/ab/libmpm_t5.so ""
b notile tools/ab/libmpm_notile.so ""
b pf tools/ab/libmpm_pf.so ""
cat gpurun_out/r2n_box.txt
