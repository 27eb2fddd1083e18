#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "c5 or rectangular_grids_3d or c2_shape" 2>&1 | tail -3
GGP_STR_WZ=16 GGP_STR_W=8 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "c5 or rectangular_grids or c2_shape or every_line" 2>&1 | tail -3
python tools/ab_bench3d.py 512 20 "default:" "wz8:GGP_STR_WZ=8" "wz16:GGP_STR_WZ=16" "wz16_wy8:GGP_STR_WZ=16,GGP_STR_W=8" "wz16_wy16:GGP_STR_WZ=16,GGP_STR_W=16" "wz8_wy8:GGP_STR_WZ=8,GGP_STR_W=8"
python tools/ab_bench3d.py 1024 6 "wz16:GGP_STR_WZ=16" "wz8:GGP_STR_WZ=8"
for lib in sb96 sb80; do echo "== C4 with libggp_$lib"; GGP_LIBRARY=$PWD/generalizedgrosspitaevskii.jl_b200/libggp_$lib.so python bench.py --workload c4 --nbatch 1024 --steps 10 --warmup 3 --no-cpu --no-extra | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chained ms/step', d['chained']['ms_per_step'], d['roofline']['per_kernel_ms'])"; done
echo "== C4 default"; python bench.py --workload c4 --nbatch 1024 --steps 10 --warmup 3 --no-cpu --no-extra | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chained ms/step', d['chained']['ms_per_step'], d['roofline']['per_kernel_ms'])"
