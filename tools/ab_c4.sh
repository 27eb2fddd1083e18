#!/bin/bash
# C4 (256^2 x NB trajectories, ComplexF64, Philox) under several libraries: tools/ab_c4.sh NB lib1.so lib2.so ...
NB=${1:-1024}; shift
for lib in default "$@"; do
  if [ "$lib" = default ]; then unset GGP_LIBRARY; else export GGP_LIBRARY=$lib; fi
  python bench.py --workload c4 --nbatch $NB --steps 20 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['per_kernel_ms']
print('$lib', 'chained ms/step %.3f'%d['chained']['ms_per_step'], 'row %.3f str %.3f'%(k['row'],k['str']), 'G pt-steps/s %.2f'%(d['chained']['value']/1e9))"
done
