#!/bin/bash
# Developer experiment: where does a tile of the SDF training kernels spend its time?  Builds variants of the library
# with the tensor work / the epilogue math compiled out and times them with scripts/bench_sdf_field.py (run on the GPU box
# after building here: the variants travel under scripts/_dbg/).
set -e
cd "$(dirname "$0")/.."
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -diag-suppress 177"
for v in NO_MMA NO_MATH SKELETON; do
  D="-DRSDF_EXP_$v"; [ $v = SKELETON ] && D="-DRSDF_EXP_NO_MMA -DRSDF_EXP_NO_MATH -DRSDF_EXP_NO_LDST"
  nvcc $FLAGS $D -c rise_sdf_b200/csrc/sdf_train.cu -o scripts/_dbg/sdf_train_$v.o
  objs=$(ls rise_sdf_b200/build/*.o | grep -v sdf_train.o)
  nvcc -shared -Wno-deprecated-gpu-targets -o scripts/_dbg/librsdf_$v.so $objs scripts/_dbg/sdf_train_$v.o
done
# the same for the inference MLP chain (csrc/mlp_fwd.cu; time with scripts/bench_mlp.py)
for v in NO_MMA NO_EPI NO_WLOAD; do
  nvcc $FLAGS -DRSDF_EXP_$v -c rise_sdf_b200/csrc/mlp_fwd.cu -o scripts/_dbg/mlp_fwd_$v.o
  objs=$(ls rise_sdf_b200/build/*.o | grep -v mlp_fwd.o)
  nvcc -shared -Wno-deprecated-gpu-targets -o scripts/_dbg/librsdf_mlp_$v.so $objs scripts/_dbg/mlp_fwd_$v.o
done
