#!/bin/bash
# build one library variant into build/lib_<name>.so: tools/build_variant.sh <name> [-DFLAG ...]
set -eu
name=$1; shift
mkdir -p build
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" \
  -I include -I belief-planning_b200/csrc belief-planning_b200/csrc/bmpc_api.cu -o build/lib_$name.so
