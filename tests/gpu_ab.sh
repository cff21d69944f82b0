#!/bin/bash
# gpu_ab.sh -- A/B timing of libpacb200.so variants (diagnostic): for each variants/*.so run the short bench and print
# ns per stereo block of k_analysis and ms per step.  Usage (on the GPU box): bash tests/gpu_ab.sh [streams] [seconds]
S=${1:-592}; SEC=${2:-10}
mkdir -p gpurun_out
cp perceptual-audio-codec_b200/libpacb200.so /tmp/lib_keep.so
for v in variants/*.so; do
  cp $v perceptual-audio-codec_b200/libpacb200.so
  python bench.py --streams $S --seconds $SEC --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('%-28s step %.2f ms  analysis %.1f ns/block  coded_bytes %d  %s' % ('$v', d['ms_per_step'], 1e6*r['avg_launch_ms']/r['blocks_per_launch'], d['config']['coded_bytes'], r['note'].split(';')[-1]))
" | tee -a gpurun_out/ab.log
done
cp /tmp/lib_keep.so perceptual-audio-codec_b200/libpacb200.so
