#!/bin/bash
# wave-fit rule of the contiguous-axis kernel (inst.cu: wave_fit_smem): off / on / forced occupancies
python - <<'PY'
import json, os, subprocess, sys
def run(args, env):
    e = dict(os.environ); e.update(env)
    out = subprocess.run([sys.executable, "bench.py"] + args + ["--no-cpu", "--no-extra"], env=e, capture_output=True, text=True).stdout
    return json.loads(out.strip().splitlines()[-1])
for wl, args in (("c2 2048", ["--steps", "200", "--warmup", "10"]), ("c3", ["--workload", "c3", "--steps", "200", "--warmup", "10"]),
                 ("c3_c64", ["--workload", "c3_c64", "--steps", "200", "--warmup", "10"])):
    for name, env in (("off", {"GGP_WAVE_FIT": "0"}), ("on", {}), ("occ6", {"GGP_WAVE_FIT": "0", "GGP_ROW_OCC": "6"}),
                      ("occ5", {"GGP_WAVE_FIT": "0", "GGP_ROW_OCC": "5"}), ("occ4", {"GGP_WAVE_FIT": "0", "GGP_ROW_OCC": "4"}),
                      ("occ3", {"GGP_WAVE_FIT": "0", "GGP_ROW_OCC": "3"})):
        try:
            d = run(args, env)
            k = d["roofline"]["per_kernel_ms"]
            print(f"{wl:8s} {name:5s} value(cold windows) {d['ms_per_step']*1e3:8.2f} us/step  chained {d['chained']['ms_per_step']*1e3:8.2f}  row {k['row']*1e3:7.2f} str {k['str']*1e3:7.2f}", flush=True)
        except Exception as ex:
            print(wl, name, "failed", ex, flush=True)
PY
