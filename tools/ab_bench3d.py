"""3-D variant of ab_bench.py: C5 workload (single GPU) under several environments.
Usage: python tools/ab_bench3d.py GRID STEPS "NAME:VAR=VAL,..." ..."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
grid, steps = sys.argv[1], sys.argv[2]
for spec in sys.argv[3:]:
    name, _, kv = spec.partition(":")
    env = dict(os.environ)
    for item in filter(None, kv.split(",")):
        k, _, v = item.partition("=")
        env[k] = v
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "c5", "--grid", grid, "--steps", steps, "--warmup", "3", "--no-cpu", "--no-extra"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        tot = d["chained"]["ms_per_step"]
        k = d["roofline"]["per_kernel_ms"]
        print(f"{grid:>5s}^3 {name:20s} cold {d['ms_per_step']:8.3f} ms  chained {tot:8.3f} ms  row {k['row']:7.3f}  str_z {k['str']:7.3f}  "
              f"str_y(each) {(tot - k['row'] - k['str']) / 2:7.3f}  frac_chained {d['roofline']['step_contract']['frac_chained']:.3f}", flush=True)
    except Exception as e:
        print(grid, name, "FAILED", repr(e), r.stderr[-400:], flush=True)
