#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_call4.log 2>&1
tail -30 gpurun_out/pytest_call4.log
for n in 1024 2048 4096 8192; do python tools/ab_bench.py $n 100 "default:"; done
python tools/ab_bench.py 4096 100 "nosephint:GGP_NO_SEP_HINT=1"
python tools/ab_bench.py 8192 30 "nosephint:GGP_NO_SEP_HINT=1"
python - <<'PY'
import json, subprocess, sys
for wl, extra in (("c3", ["--steps", "200"]), ("c4", ["--nbatch", "1024", "--steps", "10"])):
    r = subprocess.run([sys.executable, "bench.py", "--workload", wl, "--warmup", "5", "--no-cpu", "--no-extra"] + extra, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(wl, "cold ms/step", d["ms_per_step"], "chained", d["chained"]["ms_per_step"], d["roofline"]["per_kernel_ms"], "contract frac chained", d["roofline"]["step_contract"]["frac_chained"], flush=True)
    except Exception as e:
        print(wl, "FAILED", e, r.stderr[-500:])
PY
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct"
ncu $M --clock-control none -k regex:^str_kernel -s 4 -c 1 python bench.py --grid 4096 --steps 6 --warmup 3 --no-cpu --no-extra 2>&1 | grep -E "dram__|duration|hit_rate"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^str_kernel -s 4 -c 1 -f -o gpurun_out/prof_str_kernel_r01u_4096 python bench.py --grid 4096 --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_str_4096_u.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^str_kernel -s 4 -c 1 -f -o gpurun_out/prof_str_kernel_r01u_c3 python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_str_c3_u.log 2>&1
du -sh gpurun_out
