#!/bin/bash
# round-2 GPU call 15 (1 GPU): the committed 2D kernel (persistent-loop scaffolding removed): default bench line + ncu --set full
mkdir -p gpurun_out
echo start > gpurun_out/r2p_box.txt
timeout 900 python bench.py > gpurun_out/r2p_bench_c4_n1.json 2> gpurun_out/r2p_bench_c4_n1.err; echo "bench c4 rc=$?" >> gpurun_out/r2p_box.txt
CMD="python bench.py --steps 4 --warmup 3 --warm-substeps 300 --no-cpu --e2e-calls 1"
ncu --set full --clock-control none --import-source on -k regex:"k_substep2d|k_grid_tiles" -s 604 -c 4 -o gpurun_out/r2p_prof_c4 $CMD > gpurun_out/r2p_ncu_full_c4.log 2>&1
echo "ncu full c4 rc=$?" >> gpurun_out/r2p_box.txt
cat gpurun_out/r2p_box.txt
