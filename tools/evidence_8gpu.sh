#!/bin/bash
# One 8-GPU gpurun call: slab parity at world 4 and 8 (fused peer-store transposes), chunk / occupancy A-B of the slab
# pipeline at N = 8, and the full default bench line at N = 8.   TAG=r02g bash tools/evidence_8gpu.sh
TAG=${TAG:-r02g}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_slab.py -q -x -k "stores-4 or stores-8 or engines-8") > gpurun_out/pytest_slab_${TAG}.log 2>&1
tail -4 gpurun_out/pytest_slab_${TAG}.log
run() {  # name grid env...
  name=$1; grid=$2; shift 2
  env "$@" python bench.py --gpus 8 --workload c5 --grid $grid --steps 20 --warmup 3 --no-cpu --no-extra 2>gpurun_out/c5_${TAG}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=8 grid=$grid %-22s'%'$name', 'chained %.4f ms/step'%d['chained']['ms_per_step'], '%.1f G pt-steps/s'%(d['chained']['value']/1e9))"
}
if [ -n "$CHUNK_AB" ]; then
run chunks1 512 GGP_SLAB_CHUNKS=1
run chunks4 512 GGP_SLAB_CHUNKS=4
run chunks4_yocc1 512 GGP_SLAB_CHUNKS=4 GGP_SLAB_YOCC1=1
run chunks8_yocc1 512 GGP_SLAB_CHUNKS=8 GGP_SLAB_YOCC1=1
run chunks1 1024 GGP_SLAB_CHUNKS=1
run chunks4_yocc1 1024 GGP_SLAB_CHUNKS=4 GGP_SLAB_YOCC1=1
fi
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5) > gpurun_out/bench_${TAG}_n8.json 2> gpurun_out/bench_${TAG}_n8.err
tail -3 gpurun_out/bench_${TAG}_n8.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}_n8.json"))
print({k:d[k] for k in ("value","n_gpus","ms_per_step")}, d["chained"])
for k,v in d.get("sharded",{}).items():
    print(k, {a:v.get(a) for a in ("value","ms_per_step","frac_of_hbm_roofline_contract_per_gpu","efficiency_vs_n1","checksum_nk_sum","checksum_weighted_norm","error")})
PY
