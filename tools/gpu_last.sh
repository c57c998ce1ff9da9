#!/bin/bash
# tiny last GPU call: the driver tests after the exec-shim frame writer
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_drivers.py -m gpu -q > gpurun_out/r02_pytest_drivers_final.log 2>&1; echo "drivers rc=$?" > gpurun_out/r2r_box.txt
cat gpurun_out/r2r_box.txt
