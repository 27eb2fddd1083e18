#!/bin/bash
# does the nvidia-smi poller disturb the end-to-end number?  bench.py with and without it, 3 runs each
mkdir -p gpurun_out
for rep in 1 2 3; do
  for mode in clocks noclocks; do
    if [ $mode = noclocks ]; then export GGP_BENCH_NO_CLOCKS=1; else unset GGP_BENCH_NO_CLOCKS; fi
    python bench.py --steps 20 --warmup 5 --no-cpu --no-extra 2>gpurun_out/ab_e2e.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$mode rep $rep: value %.1f G  chained %.1f G  e2e %.2f G pt-steps/s'%(d['value']/1e9, d['chained']['value']/1e9, d['e2e']['value']/1e9))"
  done
done
