#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
( timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -4
python - <<'PY'
import json, subprocess, sys
for wl, extra in (("c3", ["--steps", "200"]),):
    r = subprocess.run([sys.executable, "bench.py", "--workload", wl, "--warmup", "5", "--no-cpu", "--no-extra"] + extra, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(wl, "cold ms/step", d["ms_per_step"], "chained", d["chained"]["ms_per_step"], d["roofline"]["per_kernel_ms"], "contract frac chained", d["roofline"]["step_contract"]["frac_chained"], flush=True)
    except Exception as e:
        print(wl, "FAILED", e, r.stderr[-500:])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^row_kernel -s 4 -c 1 -f -o gpurun_out/prof_row_kernel_r01x_c3 python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_row_c3_x.log 2>&1
