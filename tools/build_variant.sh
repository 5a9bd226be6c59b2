#!/bin/bash
# Experiment builds of the L=J=10 kernel next to the product library: tools/build_variant.sh NAME [-DMACRO=..]...
# -> variants/NAME.so (git-ignored, travels with gpurun); run with PROBE_SO=variants/NAME.so python tools/probe_region.py
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -DRG_ONLY_J10 "$@" \
     -shared -o variants/$name.so rg_b200/csrc/rg_b200.cu 2> variants/$name.log || { tail -20 variants/$name.log; exit 1; }
