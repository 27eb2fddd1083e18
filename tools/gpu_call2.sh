#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_save_checkpoint.py tests/test_gpu_parity.py -x -q ) > gpurun_out/pytest_call2.log 2>&1
tail -15 gpurun_out/pytest_call2.log
python tools/ab_bench.py 4096 100 "wide4:" "narrow2:GGP_STR_NARROW=1" "narrow2_notma:GGP_STR_NARROW=1,GGP_NO_TMA=1" "wide4_notma:GGP_NO_TMA=1" \
   "wide4_promo0:GGP_TMA_L2PROMO=0" "wide4_promo256:GGP_TMA_L2PROMO=256" "narrow2_promo0:GGP_STR_NARROW=1,GGP_TMA_L2PROMO=0" "wide4_notwrt:GGP_NO_TW_SMEM_RT=1" "wide4_nopdl:GGP_NO_PDL=1"
python tools/ab_bench.py 2048 300 "default:" "promo0:GGP_TMA_L2PROMO=0" "promo256:GGP_TMA_L2PROMO=256" "promo64:GGP_TMA_L2PROMO=64"
python tools/ab_bench.py 1024 500 "default:" "e8lib:GGP_LIBRARY=$PWD/generalizedgrosspitaevskii.jl_b200/libggp_e8.so" "e8lib_pdl:GGP_PDL=1,GGP_LIBRARY=$PWD/generalizedgrosspitaevskii.jl_b200/libggp_e8.so" "pdl:GGP_PDL=1"
python tools/ab_bench.py 8192 30 "default:" "notma:GGP_NO_TMA=1" "promo0:GGP_TMA_L2PROMO=0" "tpl512:GGP_LIBRARY=$PWD/generalizedgrosspitaevskii.jl_b200/libggp_e8.so" "tpl512_notma:GGP_NO_TMA=1,GGP_LIBRARY=$PWD/generalizedgrosspitaevskii.jl_b200/libggp_e8.so"
GGP_LIBRARY=$PWD/generalizedgrosspitaevskii.jl_b200/libggp_e8.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "every_line_length and (1024 or 8192)" 2>&1 | tail -3
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct"
B="python bench.py --grid 4096 --steps 6 --warmup 3 --no-cpu --no-extra"
for v in "" "GGP_STR_NARROW=1" "GGP_STR_NARROW=1 GGP_NO_TMA=1" "GGP_NO_TMA=1"; do
  echo "== dram 4096 [$v]"
  env $v ncu $M --clock-control none -k regex:^str_kernel -s 4 -c 1 $B 2>&1 | grep -E "dram__|duration|hit_rate|issue_active"
done
for k in row_kernel str_kernel; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$k -s 4 -c 1 -f -o gpurun_out/prof_${k}_r01s_8192 python bench.py --grid 8192 --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_${k}_8192.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^str_kernel -s 4 -c 1 -f -o gpurun_out/prof_str_kernel_r01s_4096 python bench.py --grid 4096 --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_str_4096_s.log 2>&1
ls gpurun_out | head -50
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^row_kernel -s 4 -c 1 -f -o gpurun_out/prof_row_kernel_r01s_c4 python bench.py --workload c4 --nbatch 512 --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_row_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^row_kernel -s 4 -c 1 -f -o gpurun_out/prof_row_kernel_r01s_c3 python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_row_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^str_kernel -s 4 -c 1 -f -o gpurun_out/prof_str_kernel_r01s_c3 python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_str_c3.log 2>&1
ls -la gpurun_out | tail -12
