#!/bin/bash
# last GPU seconds of round 2 (1 GPU, < 1.5 min): the 3D overlapped-schedule tests after the profile-span fix
mkdir -p gpurun_out
timeout 75 python -m pytest tests/test_gpu_slabs.py -m gpu -q -x -k "overlapped_schedule_3d" > gpurun_out/r02_pytest_overlap3d_last.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2u_box.txt
cat gpurun_out/r2u_box.txt
