#!/bin/bash
# last GPU call (2 GPUs, ~2 min): the overlapped two-kernel 3D slab schedule -- in-process parity test, then c5 on 2 GPUs over NCCL
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_gpu_slabs.py -m gpu -q -k "overlapped_schedule_3d or 3d_slabs" > gpurun_out/r02_pytest_overlap3d.log 2>&1; echo "test rc=$?" > gpurun_out/r2s_box.txt
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --workload c5 --steps 20 --warmup 5 --warm-substeps 400 --e2e-calls 1 > gpurun_out/r2s_c5_n2_overlap.json 2> gpurun_out/r2s_c5_n2_overlap.err; echo "c5 n2 overlap rc=$?" >> gpurun_out/r2s_box.txt
cat gpurun_out/r2s_box.txt
