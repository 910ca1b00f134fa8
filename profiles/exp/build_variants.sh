#!/bin/bash
# Builds kernel variants of librenv_b200.so into build/variants/ (git-ignored; travels with gpurun) for A/B timing
# with profiles/exp/step_variants.py --lib.  Usage: profiles/exp/build_variants.sh name "-DMACRO=.. -DMACRO=.." [...]
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
mkdir -p "$ROOT/build/variants"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --shared -Xcompiler -fPIC \
      -cudart static $flags -I "$ROOT/include" -o "$ROOT/build/variants/librenv_$name.so" \
      "$ROOT/random_envs_b200/csrc/renv_abi.cu" && echo "built $name" ) &
done
wait
