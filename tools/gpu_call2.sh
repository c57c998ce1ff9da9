#!/bin/bash
# round-2 GPU call 2: CFL re-sort controller + 3D Newton polar; before/after vs the round-1 library on the SAME moving scene; ncu
mkdir -p gpurun_out
echo start > gpurun_out/r2b_box.txt
for w in c4 c3 c2 c5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu > gpurun_out/r2b_bench_$w.json 2> gpurun_out/r2b_bench_$w.err; echo "bench $w rc=$?" >> gpurun_out/r2b_box.txt
done
MPM_LIBRARY=tools/ab/libmpm_r1.so timeout 600 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2b_bench_c4_r1lib.json 2> gpurun_out/r2b_bench_c4_r1lib.err; echo "bench c4 r1lib rc=$?" >> gpurun_out/r2b_box.txt
MPM_SKIP_HUGE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_box.txt
# ncu: launch list, then the full capture of the substep kernel (same command exited 0 directly before)
CMD="python bench.py --steps 4 --warmup 3 --warm-substeps 300 --no-cpu --e2e-calls 1"
$CMD > gpurun_out/r2b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 160 --csv --log-file gpurun_out/r2b_launches_c4.csv $CMD > gpurun_out/r2b_ncu_launch.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/r2b_box.txt
$CMD > gpurun_out/r2b_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_substep2d|k_grid_update|k_count_rank" -s 640 -c 6 -o gpurun_out/r2b_prof_c4 $CMD > gpurun_out/r2b_ncu_full.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/r2b_box.txt
cat gpurun_out/r2b_box.txt; tail -3 gpurun_out/r2b_pytest.log
