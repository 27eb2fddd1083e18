#!/bin/bash
# C5 slab: fused peer stores vs copy-engine transposes.  tools/ab_c5ce.sh NGPU GRID "name:ENV=.." ...
NG=$1; GRID=$2; shift 2
for spec in "$@"; do
  name=${spec%%:*}; kv=${spec#*:}; envs=$(echo "$kv" | tr ',' ' ')
  env $envs python bench.py --gpus $NG --workload c5 --grid $GRID --steps 20 --warmup 3 --no-cpu --no-extra 2>gpurun_out/ab_c5ce.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=$NG grid=$GRID %-26s'%'$name', 'chained %.4f ms/step'%d['chained']['ms_per_step'], '%.1f G pt-steps/s'%(d['chained']['value']/1e9))"
done
