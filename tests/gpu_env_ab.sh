#!/bin/bash
# gpu_env_ab.sh -- short bench under different experiment environment settings (diagnostic)
# usage: S=512 SEC=60 bash tests/gpu_env_ab.sh "X=0" "PAC_TILE_BLOCKS=81" ...
S=${S:-592}; SEC=${SEC:-10}
mkdir -p gpurun_out
run() {
  env "$@" python bench.py --streams $S --seconds $SEC --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%-40s step %.2f ms  analysis %.1f ns/block  %s' % ('$*', d['ms_per_step'], 1e6*r['avg_launch_ms']/r['blocks_per_launch'], r['note'].split(';')[-1]))
" | tee -a gpurun_out/env_ab.log
}
for setting in "$@"; do run $setting; done
