#!/bin/bash
# last GPU seconds of round 2 (1 GPU, ~20 s): the GPU arm of bench.py end to end on the smallest workload after the
# `config` / `engine` split of the JSON line
mkdir -p gpurun_out
timeout 30 python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu --e2e-calls 1 > gpurun_out/r02_bench_c1_last.json 2> gpurun_out/r02_bench_c1_last.err; echo "bench c1 rc=$?"
