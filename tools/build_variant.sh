#!/bin/bash
# A/B builds of libmpm.so with extra nvcc flags: tools/build_variant.sh <name> "<flags>" -> tools/ab/libmpm_<name>.so
# (git-ignored like every .so; travels to the GPU box; selected with MPM_LIBRARY=tools/ab/libmpm_<name>.so)
set -e
name=$1; flags=$2
src=$(dirname "$0")/../mpm_flip98a_b200/csrc
out=$(dirname "$0")/ab; mkdir -p $out /tmp/ab_$name
for f in mpm_engine mpm_group mpm_kernels mpm_substep2d mpm_substep3d mpm_deterministic mpm_sort; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC $flags -c $src/$f.cu -o /tmp/ab_$name/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libmpm_$name.so /tmp/ab_$name/*.o
echo built $out/libmpm_$name.so
