#!/bin/bash
# one bench workload under several libraries: tools/ab_libs.sh "<bench args>" lib1.so lib2.so ...   ("default" = in-tree)
ARGS=$1; shift
for lib in "$@"; do
  if [ "$lib" = default ]; then unset GGP_LIBRARY; else export GGP_LIBRARY=$PWD/$lib; fi
  python bench.py $ARGS --no-cpu --no-extra 2>gpurun_out/ab_libs.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['per_kernel_ms']
print('%-34s'%'$lib', 'chained %.4f ms/step'%d['chained']['ms_per_step'], 'row %.4f str %.4f'%(k['row'],k['str']), 'frac_chained %.3f'%d['roofline']['step_contract']['frac_chained'])"
done
