#!/bin/bash
# builds CE-head variants (compile-time knobs) into replay_b200/build/variants/*.so for A/B timing on the GPU box
set -e
cd "$(dirname "$0")/.."
mkdir -p replay_b200/build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --use_fast_math -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -I include"
for v in "0 0" "0 1" "4 1" "2 1" "4 0" "8 1"; do
  set -- $v
  out=replay_b200/build/variants/ce_p$1_n$2.so
  nvcc $FLAGS -DRP_CE_POLY_EVERY=$1 -DRP_CE_NBUF3=$2 -shared -o $out replay_b200/csrc/rp_ce_head.cu replay_b200/csrc/rp_host.cu -cudart static &
done
wait
ls -la replay_b200/build/variants
