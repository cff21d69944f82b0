#!/bin/bash
# gpu_e2e_ab.sh -- e2e (host buffers) leg of the default bench under different environment settings (diagnostic)
# usage: bash tests/gpu_e2e_ab.sh "X=0" "PAC_NO_MAPPED_OUT=1" ...
mkdir -p gpurun_out
for setting in "$@"; do
  env $setting python bench.py --no-cpu --no-decode --no-stages --steps 2 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-40s value %.0f (%.1f ms)  e2e %.0f (%.1f ms)  ratio %.3f' % ('$setting', d['value'], d['ms_per_step'], d['e2e']['value'], 245760e3/d['e2e']['value'], d['e2e']['value']/d['value']))
" | tee -a gpurun_out/e2e_ab.log
done
