set -x
for n in 1024 2048 4096 8192; do python bench.py --steps 200 --warmup 5 --no-cpu --no-extra --grid $n > gpurun_out/sweep_r01l_$n.json 2> gpurun_out/sweep_r01l_$n.err; python -c "
import json;d=json.load(open('gpurun_out/sweep_r01l_$n.json'));print($n,'cold',d['ms_per_step'],'chained',d['chained']['ms_per_step'],d['roofline']['per_kernel_ms'],d['roofline']['step_contract'])"; done
GGP_STR_W=4 python bench.py --steps 200 --warmup 5 --no-cpu --no-extra --grid 4096 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('W=4 4096 cold',d['ms_per_step'],'chained',d['chained']['ms_per_step'],d['roofline']['per_kernel_ms'])"
GGP_STR_W=8 python bench.py --steps 200 --warmup 5 --no-cpu --no-extra --grid 1024 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('W=8 1024 cold',d['ms_per_step'],'chained',d['chained']['ms_per_step'],d['roofline']['per_kernel_ms'])"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
ncu --set full --clock-control none --import-source on -k regex:^row_kernel -s 6 -c 1 -f -o gpurun_out/prof_row_c4_r01l python bench.py --workload c4 --nbatch 512 --steps 6 --warmup 2 --no-cpu --no-extra > gpurun_out/ncu_c4.log 2>&1
tail -3 gpurun_out/ncu_c4.log
