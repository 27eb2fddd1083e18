"""Run bench.py under several environment settings and print one compact line per variant.
Usage: python tools/ab_bench.py GRID STEPS "NAME:VAR=VAL,VAR=VAL" ...   (NAME: alone = default environment)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
grid, steps = sys.argv[1], sys.argv[2]
for spec in sys.argv[3:]:
    name, _, kv = spec.partition(":")
    env = dict(os.environ)
    for item in filter(None, kv.split(",")):
        k, _, v = item.partition("=")
        env[k] = v
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--grid", grid, "--steps", steps, "--warmup", "5", "--no-cpu", "--no-extra"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        k = d["roofline"]["per_kernel_ms"]
        print(f"{grid:>5s} {name:28s} cold {1e3*d['ms_per_step']:8.2f} us  chained {1e3*d['chained']['ms_per_step']:8.2f} us  "
              f"row {1e3*k['row']:7.2f}  str {1e3*k['str']:7.2f}  frac_chained {d['roofline']['step_contract']['frac_chained']:.3f}  "
              f"e2e {d['e2e']['value']/1e9:6.1f} G", flush=True)
    except Exception as e:
        print(grid, name, "FAILED", repr(e), r.stderr[-400:], flush=True)
