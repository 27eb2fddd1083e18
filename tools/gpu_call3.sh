#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_call3.log 2>&1
tail -30 gpurun_out/pytest_call3.log
for n in 1024 2048 4096 8192; do python tools/ab_bench.py $n 100 "default:"; done
python tools/ab_bench.py 4096 100 "wide:GGP_STR_WIDE=1" "notma:GGP_NO_TMA=1"
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct,l1tex__t_sector_hit_rate.pct"
ncu $M --clock-control none -k regex:^str_kernel -s 4 -c 1 python bench.py --grid 4096 --steps 6 --warmup 3 --no-cpu --no-extra 2>&1 | grep -E "dram__|duration|hit_rate|issue_active"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^str_kernel -s 4 -c 1 -f -o gpurun_out/prof_str_kernel_r01t_4096 python bench.py --grid 4096 --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_str_4096_t.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^row_kernel -s 4 -c 1 -f -o gpurun_out/prof_row_kernel_r01t_c4 python bench.py --workload c4 --nbatch 512 --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_row_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^row_kernel -s 4 -c 1 -f -o gpurun_out/prof_row_kernel_r01t_c3 python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_row_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^str_kernel -s 4 -c 1 -f -o gpurun_out/prof_str_kernel_r01t_c3 python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_str_c3.log 2>&1
ls -la gpurun_out | tail -12; du -sh gpurun_out
