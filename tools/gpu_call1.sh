#!/bin/bash
# round-2 GPU call 1: full GPU test suite (per file, own process), then the bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_box.txt 2>&1
free -g >> gpurun_out/r2_box.txt; nproc >> gpurun_out/r2_box.txt
for f in test_gpu_parity test_gpu_slabs test_drivers test_gpu_fullsize; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q -s --durations=8 > gpurun_out/r2_$f.log 2>&1
  echo "$f rc=$?" >> gpurun_out/r2_box.txt
done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_box.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; echo "bench c4 rc=$?" >> gpurun_out/r2_box.txt
for w in c3 c2 c5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; echo "bench $w rc=$?" >> gpurun_out/r2_box.txt
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "bench ref rc=$?" >> gpurun_out/r2_box.txt
cat gpurun_out/r2_box.txt
tail -5 gpurun_out/r2_test_gpu_*.log
