#!/bin/bash
# round-2 GPU call 4: cell-ordered storage, deterministic mode, slab protocol fix, 3D packed kernels; ncu of the substep kernel
mkdir -p gpurun_out
echo start > gpurun_out/r2d_box.txt
for f in test_gpu_slabs test_gpu_deterministic test_gpu_parity test_gpu_fullsize test_drivers; do
  MPM_SKIP_HUGE=1 timeout 1200 python -m pytest tests/$f.py -m gpu -q --durations=5 > gpurun_out/r2d_$f.log 2>&1
  echo "$f rc=$?" >> gpurun_out/r2d_box.txt
done
for w in c4 c2 c3 c5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu > gpurun_out/r2d_bench_$w.json 2> gpurun_out/r2d_bench_$w.err; echo "bench $w rc=$?" >> gpurun_out/r2d_box.txt
done
CMD="python bench.py --steps 4 --warmup 3 --warm-substeps 300 --no-cpu --e2e-calls 1"
$CMD > gpurun_out/r2d_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_substep2d" -s 330 -c 2 -o gpurun_out/r2d_prof_c4 $CMD > gpurun_out/r2d_ncu_full.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/r2d_box.txt
CMD5="python bench.py --workload c5 --steps 4 --warmup 3 --warm-substeps 200 --no-cpu --e2e-calls 1"
$CMD5 > gpurun_out/r2d_plain5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_p2g_cells|k_g2p_naive" -s 500 -c 2 -o gpurun_out/r2d_prof_c5 $CMD5 > gpurun_out/r2d_ncu_full5.log 2>&1
echo "ncu c5 rc=$?" >> gpurun_out/r2d_box.txt
cat gpurun_out/r2d_box.txt; tail -3 gpurun_out/r2d_test_gpu_*.log
