set -x
python bench.py --workload c4 --nbatch 512 --steps 20 --warmup 3 --no-cpu --no-extra | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('c4 cold',d['ms_per_step'],'chained',d['chained']['ms_per_step'],d['roofline']['per_kernel_ms'])"
GGP_NO_PUMP_CONST=1 python bench.py --workload c4 --nbatch 512 --steps 20 --warmup 3 --no-cpu --no-extra | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('c4 no pump const cold',d['ms_per_step'],'chained',d['chained']['ms_per_step'],d['roofline']['per_kernel_ms'])"
python bench.py --workload c3 --steps 100 --warmup 3 --no-cpu --no-extra | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('c3 cold',d['ms_per_step'],'chained',d['chained']['ms_per_step'],d['roofline']['per_kernel_ms'])"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:^row_kernel -s 6 -c 1 -f -o gpurun_out/prof_row_c4_r01q python bench.py --workload c4 --nbatch 512 --steps 6 --warmup 2 --no-cpu --no-extra > gpurun_out/ncu_c4.log 2>&1
