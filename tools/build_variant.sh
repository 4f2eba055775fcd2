#!/bin/bash
# usage: tools/build_variant.sh <name> [-DFLAG=VALUE ...]   -> variants/lib<name>.so (development: kernel variants side by side,
# selected at run time with ASP_B200_LIBRARY=variants/lib<name>.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
src=annealing-sign-problem_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" -I include -o variants/lib$name.so \
  $src/operator.cu $src/extract.cu $src/extract_fused.cu $src/legacy.cu $src/reduce.cu $src/anneal.cu $src/greedy.cu $src/apply.cu $src/host.cu $src/peer.cu $src/sampling.cu
echo built variants/lib$name.so
