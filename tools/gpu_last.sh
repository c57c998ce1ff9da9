#!/bin/bash
# last GPU seconds of round 2 (1 GPU, < 1 min): smoke() and the 2D overlapped-schedule tests on the final libmpm.so
mkdir -p gpurun_out
timeout 18 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_last.log 2>&1; echo "smoke rc=$?" > gpurun_out/r2v_box.txt
timeout 35 python -m pytest tests/test_gpu_slabs.py -m gpu -q -x -k "overlapped_schedule_wide" > gpurun_out/r02_pytest_overlap2d_last.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_box.txt
cat gpurun_out/r2v_box.txt
