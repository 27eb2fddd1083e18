#!/bin/bash
# gpurun --gpus 2: slab-decomposition tests, bench.py at N = 2 (C2 replicas + sharded C4 + C5 512^3 slab), C5 1024^3 at N = 2
# 2 GPUs: slab decomposition tests (both transpose paths), sharded-ensemble invariance, bench at N = 2
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_slab.py -x -q ) 2>&1 | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 500 --warmup 5 > gpurun_out/bench_${TAG:-r01}_n2.json 2> gpurun_out/bench_${TAG:-r01}_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG:-r01}_n2.json").read().strip().splitlines()[-1])
    print("N=2 C2 replicas cold G", d["value"]/1e9, "chained", d["chained"]["value"]/1e9)
    for k,v in d["extra"].items(): print(k, {a:b for a,b in v.items() if a in ("value","ms_per_step","per_kernel_ms","nk_sum","error","frac_of_hbm_roofline_contract")})
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_${TAG:-r01}_n2.err").read()[-2000:])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c5 --grid 1024 --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_${TAG:-r01}_c5_1024_n2.json 2> gpurun_out/bench_${TAG:-r01}_c5_1024_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG:-r01}_c5_1024_n2.json").read().strip().splitlines()[-1])
    print("N=2 C5 1024^3: G pt-steps/s", d["chained"]["value"]/1e9, "ms/step", d["chained"]["ms_per_step"], d["roofline"]["per_kernel_ms"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_${TAG:-r01}_c5_1024_n2.err").read()[-2000:])
PY
