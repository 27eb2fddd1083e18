#!/bin/bash
# C5 slab runs under several GGP_SLAB_CHUNKS settings: tools/ab_c5.sh NGPU GRID chunks...
NG=$1; GRID=$2; shift 2
for ch in "$@"; do
  GGP_SLAB_CHUNKS=$ch python bench.py --gpus $NG --workload c5 --grid $GRID --steps 20 --warmup 3 --no-cpu --no-extra 2>gpurun_out/ab_c5.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['per_kernel_ms']
print('N=$NG grid=$GRID chunks=$ch', 'chained ms/step %.3f'%d['chained']['ms_per_step'], 'G pt-steps/s %.1f'%(d['chained']['value']/1e9), 'per-kernel(unchunked)', {a:round(b,3) for a,b in k.items()})"
done
