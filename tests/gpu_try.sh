#!/bin/bash
# gpu_try.sh -- run the short and the full bench with the in-tree library, keeping stderr (diagnostic)
mkdir -p gpurun_out
for cfg in "592 10" "4096 60"; do
  set -- $cfg
  python bench.py --streams $1 --seconds $2 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/try_$1.json 2> gpurun_out/try_$1.err
  echo "bench $1x$2 rc=$?"; tail -c 600 gpurun_out/try_$1.json | head -c 600; echo; tail -5 gpurun_out/try_$1.err | cut -c1-400
done
