#!/bin/bash
# round-2 evidence run (1 GPU): full GPU test suite, smoke, every bench line, ncu launch list + --set full captures
mkdir -p gpurun_out
echo start > gpurun_out/r02_box.txt
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,power.limit --format=csv >> gpurun_out/r02_box.txt
timeout 2700 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" >> gpurun_out/r02_box.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_box.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "reference arm rc=$?" >> gpurun_out/r02_box.txt
timeout 900 python bench.py > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/r02_bench_c4_n1.err; echo "bench c4 rc=$?" >> gpurun_out/r02_box.txt
for w in c1 c2 c3 c5; do
  timeout 600 python bench.py --workload $w --steps 40 --warmup 5 --no-cpu > gpurun_out/r02_bench_${w}_n1.json 2> gpurun_out/r02_bench_${w}_n1.err; echo "bench $w rc=$?" >> gpurun_out/r02_box.txt
done
CMD="python bench.py --steps 4 --warmup 3 --warm-substeps 300 --no-cpu --e2e-calls 1"
$CMD > gpurun_out/r02_plain_c4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file gpurun_out/r02_launches_c4.csv $CMD > gpurun_out/r02_ncu_launch.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/r02_box.txt
ncu --set full --clock-control none --import-source on -k regex:"k_substep2d|k_grid_tiles" -s 604 -c 4 -o gpurun_out/r02_prof_c4 $CMD > gpurun_out/r02_ncu_full_c4.log 2>&1
echo "ncu full c4 rc=$?" >> gpurun_out/r02_box.txt
CMD5="python bench.py --workload c5 --steps 4 --warmup 3 --warm-substeps 200 --no-cpu --e2e-calls 1"
$CMD5 > gpurun_out/r02_plain_c5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_p2g_cells|k_g2p3" -s 404 -c 2 -o gpurun_out/r02_prof_c5 $CMD5 > gpurun_out/r02_ncu_full_c5.log 2>&1
echo "ncu full c5 rc=$?" >> gpurun_out/r02_box.txt
cat gpurun_out/r02_box.txt
