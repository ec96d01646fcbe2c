#!/bin/bash
# A/B of an environment knob inside ONE gpurun call (for two BUILDS of the library: RP_B200_LIB=<path to the other .so> vs "")
# A/B of an environment knob inside ONE gpurun call (box-to-box variance is ~2-3 %): tools/ab_env.sh VAR "v1 v2 v1 v2" [bench args]
VAR=$1; VALS=$2; shift 2
for v in $VALS; do
  env $VAR=$v timeout 250 python bench.py --steps 50 --warmup 3 --no-cpu --no-device-batches "$@" 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d.get('scoring') or {}; print('$VAR=$v', round(d['ms_per_step'],4), 'ms/step', round(s.get('ms_per_call',0),4), 'ms/predict call')"
done
