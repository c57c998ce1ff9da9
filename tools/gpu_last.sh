#!/bin/bash
# last GPU call of round 2 (1 GPU): the GPU test suite on the final tree (the 245 M-particle case ran in the evidence call)
mkdir -p gpurun_out
MPM_SKIP_HUGE=1 timeout 500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2q_box.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2q_box.txt
cat gpurun_out/r2q_box.txt
