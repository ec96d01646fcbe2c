#!/bin/bash
# run on the GPU box (under gpurun): launch list of one training step + full captures of the dominant kernels
set -x
mkdir -p gpurun_out
export RP_PROFILE=1
ncu --metrics gpu__time_duration.sum --clock-control none -s 480 -c 110 --csv --log-file gpurun_out/launches_r1c.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu --no-scoring > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ce_bwd_kernel" -s 4 -c 2 -o gpurun_out/prof_ce_r1b \
    python tools/run_ce_once.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"score_topk_kernel" -s 1 -c 1 -o gpurun_out/prof_topk_r1c \
    python tools/run_topk_once.py 4096 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"gemm_kernel|attn_fwd_kernel|attn_softmax_bwd|layernorm" -s 500 -c 16 -o gpurun_out/prof_body_r1a \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu --no-scoring > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
