#!/bin/bash
# Experiment builds of one Taylor order next to the product library: tools/build_variant.sh NAME [J=10] [-DMACRO=..]...
# -> variants/NAME.so (git-ignored, travels with gpurun); run with PROBE_SO=variants/NAME.so python tools/probe_region.py
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
name=$1; shift
J=10
if [[ "$1" =~ ^J=([0-9]+)$ ]]; then J=${BASH_REMATCH[1]}; shift; fi
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC"
nvcc $FLAGS "$@" -c -o variants/${name}_host.o rg_b200/csrc/rg_b200.cu 2> variants/$name.log || { tail -20 variants/$name.log; exit 1; }
for big in 0 1; do
  nvcc $FLAGS -Xptxas -v "$@" -DRG_INST_J=$J -DRG_INST_BIG=$big -c -o variants/${name}_inst$big.o rg_b200/csrc/rg_kernel_inst.cu 2>> variants/$name.log || { tail -20 variants/$name.log; exit 1; }
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/$name.so variants/${name}_host.o variants/${name}_inst0.o variants/${name}_inst1.o
rm -f variants/${name}_host.o variants/${name}_inst?.o
echo "variants/$name.so (J=$J only)"
