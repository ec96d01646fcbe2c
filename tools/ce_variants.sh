#!/bin/bash
# builds CE-head variants (compile-time knobs) into replay_b200/build/variants/*.so for A/B timing on the GPU box
set -e
cd "$(dirname "$0")/.."
rm -rf replay_b200/build/variants; mkdir -p replay_b200/build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --use_fast_math -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -I include"
# columns: A_TMEM  column groups  poly-every  ablate (0 = product, 1 = no exponentials, 2 = no epilogue work)  ring depth  issue order (RP_CE_ORDER)  epilogue warp sets (RP_CE_GROUPS)
if [ -n "$VARIANTS" ]; then eval "set -- $VARIANTS"; else set -- "1 2 8 0 4 0" "1 2 8 0 4 1" "1 2 8 0 4 2" "1 2 8 2 4 1" "1 2 8 0 3 1" "1 2 0 0 4 1" "1 2 4 0 4 1"; fi
for v in "$@"; do
  set -- $v
  out=replay_b200/build/variants/ce_atmem$1_cg$2_poly$3_abl$4_st$5_ord${6:-1}_grp${7:-2}.so
  nvcc $FLAGS -DRP_CE_A_TMEM=$1 -DRP_CE_BWD_CG=$2 -DRP_CE_POLY_EVERY_BWD=$3 -DRP_CE_ABLATE=$4 -DRP_CE_NSTAGE_D128=$5 -DRP_CE_ORDER=${6:-1} -DRP_CE_GROUPS=${7:-2} -shared -o $out \
    replay_b200/csrc/rp_ce_head.cu replay_b200/csrc/rp_gemm.cu replay_b200/csrc/rp_host.cu -cudart static &
done
wait
ls replay_b200/build/variants
