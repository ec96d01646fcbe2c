#!/bin/bash
# run on the GPU box (under gpurun): launch list of the training step + full captures of the kernels added this round
set -x
mkdir -p gpurun_out
TAG=${1:-r1d}
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 450 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --batch 256 --no-graph --no-cpu --no-scoring --no-device-batches > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_kernel|attn_fwd_kernel" -s 8 -c 2 -o gpurun_out/prof_attn_${TAG} \
    python bench.py --steps 2 --warmup 3 --batch 256 --no-graph --no-cpu --no-scoring --no-device-batches > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ce_bwd_kernel" -s 4 -c 2 -o gpurun_out/prof_ce_${TAG} \
    python tools/run_ce_once.py > /dev/null 2>&1
ls -la gpurun_out/
