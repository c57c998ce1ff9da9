#!/bin/bash
# round-2 GPU call 6 (2 GPUs): slab protocol v2 over NCCL, overlap default, c5 / strong scaling, mpm_group over 2 real devices
mkdir -p gpurun_out
echo start > gpurun_out/r2g_box.txt
nvidia-smi --query-gpu=index,name --format=csv >> gpurun_out/r2g_box.txt
run() { # name, args...
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 "$@" > gpurun_out/r2g_$name.json 2> gpurun_out/r2g_$name.err
  rc=$?
  echo "$name rc=$rc" >> gpurun_out/r2g_box.txt
  return $rc
}
run c4_n2 || run c4_n2_nooverlap --no-overlap
run c5_n2 --workload c5
run c4_n2_strong --scaling strong --warm-substeps 1000
timeout 300 python - > gpurun_out/r2g_group.log 2>&1 <<'PY'
import numpy as np, time
import mpm_flip98a_b200 as mpm
from mpm_flip98a_b200 import scenes
n=2048; dt,vol=scenes.scaled_constants(n)
p=scenes.slab_fill_2d(n,per_side=3,swirl=3.0)
for devs in ([0],[0,1]):
    with mpm.Group(devs, dim=2, n_grid=n, capacity=len(p), dt=dt, vol_p=vol) as g:
        g.upload(p); g.substep(200); out=g.read(); assert g.poll_status()==0
        t=time.time(); g.substep(200); g.synchronize(); el=time.time()-t; out=g.read()
        print(devs, 'slabs', g.slabs(), 'particle-substeps/s %.3e'%(len(p)*200/el), 'com', out[:,0:2].mean(0), 'ke', 0.5*(out[:,2:4].astype(np.float64)**2).sum())
PY
echo "group rc=$?" >> gpurun_out/r2g_box.txt
cat gpurun_out/r2g_box.txt
