#!/bin/bash
# round-2 GPU call 5a: touched-tile grid pass + deterministic fix + c1; all GPU suites except the 245M case
mkdir -p gpurun_out
echo start > gpurun_out/r2f_box.txt
for f in test_gpu_deterministic test_gpu_slabs test_gpu_parity test_gpu_fullsize test_drivers; do
  MPM_SKIP_HUGE=1 timeout 1200 python -m pytest tests/$f.py -m gpu -q --durations=3 > gpurun_out/r2f_$f.log 2>&1
  echo "$f rc=$?" >> gpurun_out/r2f_box.txt
done
for w in c4 c2 c3 c1; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu > gpurun_out/r2f_bench_$w.json 2> gpurun_out/r2f_bench_$w.err; echo "bench $w rc=$?" >> gpurun_out/r2f_box.txt
done
cat gpurun_out/r2f_box.txt; tail -3 gpurun_out/r2f_test_*.log
