set -x
timeout 600 python -m pytest tests/test_gpu_slab.py -x -q 2>&1 | tail -8
for mode in p2p nccl; do
  if [ $mode = nccl ]; then export GGP_SLAB_NCCL=1; else unset GGP_SLAB_NCCL; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c5 --grid 512 --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/c5_${mode}_n2.json 2> gpurun_out/c5_${mode}_n2.err
  python -c "
import json;d=json.load(open('gpurun_out/c5_${mode}_n2.json'));print('$mode',d['value'],d['ms_per_step'],d.get('chained'),d['roofline'].get('per_kernel_ms'))"
  tail -3 gpurun_out/c5_${mode}_n2.err
done
