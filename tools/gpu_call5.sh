#!/bin/bash
# round-2 GPU call 5 (2 GPUs): slab protocol v2 over NCCL, overlap default, c5 / strong scaling, mpm_group over 2 real devices
mkdir -p gpurun_out
echo start > gpurun_out/r2e_box.txt
nvidia-smi --query-gpu=index,name --format=csv >> gpurun_out/r2e_box.txt
run() { # name, args...
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 "$@" > gpurun_out/r2e_$name.json 2> gpurun_out/r2e_$name.err
  echo "$name rc=$?" >> gpurun_out/r2e_box.txt
}
run c4_n2
run c4_n2_nooverlap --no-overlap
run c4_n2_strong --scaling strong
run c5_n2 --workload c5
run c5_n2_strong --workload c5 --scaling strong
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2e_c4_n1.json 2> gpurun_out/r2e_c4_n1.err; echo "c4_n1 rc=$?" >> gpurun_out/r2e_box.txt
timeout 300 python - > gpurun_out/r2e_group.log 2>&1 <<'PY'
import numpy as np, time
import mpm_flip98a_b200 as mpm
from mpm_flip98a_b200 import scenes
n=2048; dt,vol=scenes.scaled_constants(n)
p=scenes.slab_fill_2d(n,per_side=3,swirl=3.0)
for devs in ([0],[0,1]):
    with mpm.Group(devs, dim=2, n_grid=n, capacity=len(p), dt=dt, vol_p=vol) as g:
        g.upload(p); g.substep(200); out=g.read(); assert g.poll_status()==0
        t=time.time(); g.substep(200); out=g.read(); el=time.time()-t
        print(devs, 'slabs', g.slabs(), 'particle-substeps/s incl. read %.3e'%(len(p)*200/el), 'com', out[:,0:2].mean(0), 'ke', 0.5*(out[:,2:4].astype(np.float64)**2).sum())
PY
echo "group rc=$?" >> gpurun_out/r2e_box.txt
cat gpurun_out/r2e_box.txt
