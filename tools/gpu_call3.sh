#!/bin/bash
# round-2 GPU call 3: slab protocol v2 + group handle + chunk work list + 3D Newton polar (tests), benches, occupancy A/B
mkdir -p gpurun_out
echo start > gpurun_out/r2c_box.txt
for f in test_gpu_slabs test_gpu_parity test_drivers test_gpu_fullsize; do
  MPM_SKIP_HUGE=1 timeout 1200 python -m pytest tests/$f.py -m gpu -q --durations=5 > gpurun_out/r2c_$f.log 2>&1
  echo "$f rc=$?" >> gpurun_out/r2c_box.txt
done
for w in c4 c2 c3 c5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu > gpurun_out/r2c_bench_$w.json 2> gpurun_out/r2c_bench_$w.err; echo "bench $w rc=$?" >> gpurun_out/r2c_box.txt
done
for v in m8c704 m9c640 m6c768; do
  MPM_LIBRARY=tools/ab/libmpm_$v.so timeout 600 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu --warm-substeps 800 --e2e-calls 1 > gpurun_out/r2c_bench_c4_$v.json 2> gpurun_out/r2c_bench_c4_$v.err; echo "bench c4 $v rc=$?" >> gpurun_out/r2c_box.txt
done
timeout 600 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu --warm-substeps 800 --e2e-calls 1 > gpurun_out/r2c_bench_c4_base800.json 2> gpurun_out/r2c_bench_c4_base800.err; echo "bench c4 base800 rc=$?" >> gpurun_out/r2c_box.txt
cat gpurun_out/r2c_box.txt; tail -3 gpurun_out/r2c_test_gpu_*.log
