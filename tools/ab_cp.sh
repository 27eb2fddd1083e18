#!/bin/bash
# two-component problems: component-parallel kernels (default, and with an 80-register budget) vs the
# one-thread-both-components kernels (GGP_NO_CP=1); then the parity tests that cover two components
for wl in c3 c3_c64; do
  bash tools/ab_wl.sh "--workload $wl --steps 200 --warmup 10" "$wl-cp64:GGP_X=0" "$wl-cp80:GGP_LIBRARY=$PWD/gpurun_ship/libggp_cp80.so" "$wl-nocp:GGP_NO_CP=1"
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_baseline_sizes.py tests/test_gpu_dense_pump.py tests/test_gpu_generic.py -m gpu -q -x -k "kind or exciton or c3 or two or dense or fused or wrapper or known" 2>&1 | tail -3
