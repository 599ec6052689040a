#!/bin/bash
# Builds an experiment variant of the library into tools/_abl/ (git-ignored; travels to the GPU box with gpurun).
# usage: tools/build_variant.sh <name> [-DFLAG=VALUE ...]   ->  tools/_abl/libtalfe_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p tools/_abl
nvcc -O3 -std=c++17 -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -lineinfo "$@" \
    tal_asrd_b200/csrc/talfe.cu -o tools/_abl/libtalfe_${name}.so -ldl
echo tools/_abl/libtalfe_${name}.so
