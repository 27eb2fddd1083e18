#!/bin/bash
# any bench workload under several environment settings: tools/ab_wl.sh "<bench args>" "name:VAR=VAL,VAR=VAL" ...
ARGS=$1; shift
for spec in "$@"; do
  name=${spec%%:*}; kv=${spec#*:}
  envs=$(echo "$kv" | tr ',' ' ')
  env $envs python bench.py $ARGS --no-cpu --no-extra 2>gpurun_out/ab_wl.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['per_kernel_ms']
print('%-28s'%'$name', 'chained %.4f ms/step'%d['chained']['ms_per_step'], 'row %.4f str %.4f'%(k['row'],k['str']), 'frac_chained %.3f'%d['roofline']['step_contract']['frac_chained'])"
done
