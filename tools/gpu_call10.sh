#!/bin/bash
mkdir -p gpurun_out
run() { # name grid steps env...
  name=$1; grid=$2; steps=$3; shift 3
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload c5 --grid $grid --steps $steps --warmup 3 --no-cpu --no-extra > gpurun_out/c5_$name.json 2> gpurun_out/c5_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c5_$name.json").read().strip().splitlines()[-1])
    tot=d["chained"]["ms_per_step"]; k=d["roofline"]["per_kernel_ms"]
    print("$name grid $grid: ms/step", round(tot,3), "G pt-steps/s", round(d["chained"]["value"]/1e9,1), "row", round(k["row"],3), "str_z", round(k["str"],3), "str_y each", round((tot-k["row"]-k["str"])/2,3))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/c5_$name.err").read()[-800:])
PY
}
run default512 512 20 A=1
run w8_512 512 20 GGP_STR_W=8 GGP_STR_WZ=8
run w16_512 512 20 GGP_STR_W=16 GGP_STR_WZ=16
run w16y_512 512 20 GGP_STR_W=16
run w16_1024 1024 6 GGP_STR_W=16 GGP_STR_WZ=16
