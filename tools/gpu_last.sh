#!/bin/bash
# last GPU call of round 2 (1 GPU, < 3 min): the GPU suite without the three full-size cases (covered by
# profiles/r02_pytest_gpu.log / r02_pytest_gpu_final.log) and smoke(), on the final tree
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q -x --ignore=tests/test_gpu_fullsize.py --durations=5 > gpurun_out/r02_pytest_gpu_last.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2t_box.txt
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_last.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2t_box.txt
cat gpurun_out/r2t_box.txt
