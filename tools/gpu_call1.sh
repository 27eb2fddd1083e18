#!/bin/bash
# round-1 session-4 evidence refresh: GPU tests, headline bench, ncu launch list + full captures, size sweep
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q --durations=12 ) > gpurun_out/pytest_gpu_r01r.log 2>&1
tail -25 gpurun_out/pytest_gpu_r01r.log
python ncu_capture.py r01r --kernels row_kernel,str_kernel
for n in 1024 4096 8192; do
  timeout 300 python bench.py --grid $n --steps 200 --warmup 5 --no-cpu --no-extra > gpurun_out/bench_r01r_$n.json 2> gpurun_out/bench_r01r_$n.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_r01r_$n.json"))
print($n, "cold ms/step", d["ms_per_step"], "chained", d["chained"]["ms_per_step"], d["roofline"]["per_kernel_ms"], d["roofline"]["step_contract"])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^str_kernel -s 8 -c 1 -f -o gpurun_out/prof_str_kernel_r01r_4096 python bench.py --grid 4096 --steps 12 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_str_4096.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^row_kernel -s 8 -c 1 -f -o gpurun_out/prof_row_kernel_r01r_4096 python bench.py --grid 4096 --steps 12 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_row_4096.log 2>&1
ls -la gpurun_out
