#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
python tools/ab_check.py GGP_STR_PF 4096 complex64 3
python tools/ab_bench.py 4096 100 "default:" "pf0:GGP_STR_PF=0" "pf148:GGP_STR_PF=148" "pf592:GGP_STR_PF=592" "wide:GGP_STR_WIDE=1" "wide_pf0:GGP_STR_WIDE=1,GGP_STR_PF=0"
python tools/ab_bench.py 8192 30 "default:" "pf0:GGP_STR_PF=0" "pf296:GGP_STR_PF=296"
python tools/ab_bench.py 2048 300 "default:" "wide:GGP_STR_WIDE=1" "pf296:GGP_STR_PF=296"
free -g | head -2; nproc
MEMG=$(free -g | awk '/Mem:/{print $7}')
if [ "$MEMG" -gt 150 ]; then
  ( time timeout 900 python bench.py --workload c5 --grid 1024 --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_c5_1024_n1.json 2> gpurun_out/bench_c5_1024_n1.err ) 2>&1 | tail -3
  python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_c5_1024_n1.json").read().strip().splitlines()[-1])
    print("C5 1024^3 N=1: cold ms/step", d["ms_per_step"], "chained", d["chained"]["ms_per_step"], d["roofline"]["per_kernel_ms"], d["roofline"]["step_contract"])
except Exception as e:
    print("c5 failed", e); print(open("gpurun_out/bench_c5_1024_n1.err").read()[-1500:])
PY
fi
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^str_kernel -s 4 -c 1 -f -o gpurun_out/prof_str_kernel_r01v_4096 python bench.py --grid 4096 --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_str_4096_v.log 2>&1
du -sh gpurun_out
