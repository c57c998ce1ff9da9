#!/bin/bash
# round-2 GPU call 10 (1 GPU): persistent 2D substep kernel A/B, ncu of the 3D kernels
mkdir -p gpurun_out
echo start > gpurun_out/r2k_box.txt
MPM_SKIP_HUGE=1 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py -m gpu -q -x --durations=3 > gpurun_out/r2k_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2k_box.txt
for w in c4 c2 c3; do
for v in default nopersist; do
  lib=tools/ab/libmpm_$v.so; [ $v = default ] && lib=mpm_flip98a_b200/libmpm.so
  MPM_LIBRARY=$lib timeout 600 python bench.py --workload $w --steps 40 --warmup 5 --no-cpu --e2e-calls 1 > gpurun_out/r2k_bench_${w}_$v.json 2> gpurun_out/r2k_bench_${w}_$v.err; echo "bench $w $v rc=$?" >> gpurun_out/r2k_box.txt
done
done
CMD5="python bench.py --workload c5 --steps 4 --warmup 3 --warm-substeps 200 --no-cpu --e2e-calls 1"
ncu --set full --clock-control none --import-source on -k regex:"k_p2g_cells|k_g2p3" -s 500 -c 2 -o gpurun_out/r2k_prof_c5 $CMD5 > gpurun_out/r2k_ncu_full5.log 2>&1
echo "ncu c5 rc=$?" >> gpurun_out/r2k_box.txt
cat gpurun_out/r2k_box.txt
