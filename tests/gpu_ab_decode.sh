#!/bin/bash
# gpu_ab_decode.sh -- A/B of libpacb200.so variants on the decode-only leg (diagnostic)
S=${1:-1024}; SEC=${2:-30}
mkdir -p gpurun_out
cp perceptual-audio-codec_b200/libpacb200.so /tmp/lib_keep.so
for v in variants/*.so; do
  cp $v perceptual-audio-codec_b200/libpacb200.so
  python bench.py --streams $S --seconds $SEC --steps 3 --warmup 2 --no-cpu --no-e2e --decode 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); x=d['decode']
print('%-28s decode %.0f audio-s/s  %.2f ms/step  %s' % ('$v', x['value'], x['ms_per_step'], {k: round(v/3,2) for k,v in x['kernels_ms'].items()}))
" | tee -a gpurun_out/ab_decode.log
done
cp /tmp/lib_keep.so perceptual-audio-codec_b200/libpacb200.so
