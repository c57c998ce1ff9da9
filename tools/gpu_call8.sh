#!/bin/bash
# round-2 GPU call 8 (1 GPU): new 3D snow projection (polar + symmetric Jacobi): parity suites, c5 A/B unfused vs fused
mkdir -p gpurun_out
echo start > gpurun_out/r2i_box.txt
for f in test_gpu_parity test_gpu_deterministic test_gpu_slabs test_gpu_fullsize; do
  MPM_SKIP_HUGE=1 timeout 1200 python -m pytest tests/$f.py -m gpu -q -x --durations=3 > gpurun_out/r2i_$f.log 2>&1
  echo "$f rc=$?" >> gpurun_out/r2i_box.txt
done
timeout 600 python bench.py --workload c5 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2i_bench_c5.json 2> gpurun_out/r2i_bench_c5.err; echo "bench c5 rc=$?" >> gpurun_out/r2i_box.txt
MPM_LIBRARY=tools/ab/libmpm_fuse3d.so timeout 600 python bench.py --workload c5 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2i_bench_c5_fuse3d.json 2> gpurun_out/r2i_bench_c5_fuse3d.err; echo "bench c5 fuse3d rc=$?" >> gpurun_out/r2i_box.txt
cat gpurun_out/r2i_box.txt
