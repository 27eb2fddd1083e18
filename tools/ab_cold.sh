#!/bin/bash
# headline (cold step windows) under several environment settings: tools/ab_cold.sh "<bench args>" "name:VAR=VAL,VAR=VAL" ...
ARGS=$1; shift
for spec in "$@"; do
  name=${spec%%:*}; kv=${spec#*:}
  envs=$(echo "$kv" | tr ',' ' ')
  env $envs python bench.py $ARGS --no-cpu --no-extra 2>gpurun_out/ab_cold.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['per_kernel_ms']
print('%-20s'%'$name', 'value(cold windows) %.2f us/step  frac_cold %.3f | chained %.2f | cold kernels row %.2f str %.2f'%(d['ms_per_step']*1e3, d['roofline']['step_contract']['frac_cold'], d['chained']['ms_per_step']*1e3, k['row']*1e3, k['str']*1e3))"
done
