#!/bin/bash
# One gpurun call that produces the evidence kept under profiles/: GPU test log, headline bench line, ncu launch list,
# ncu --set full captures of both hot kernels (2048^2, 1024^2, 8192^2), size sweep.  TAG=r01x bash tools/evidence_1gpu.sh
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/pytest_gpu_${TAG:-r01}.log 2>&1
tail -14 gpurun_out/pytest_gpu_${TAG:-r01}.log
python ncu_capture.py ${TAG:-r01} --kernels row_kernel,str_kernel
for n in 1024 4096 8192; do
  timeout 300 python bench.py --grid $n --steps 200 --warmup 5 --no-cpu --no-extra > gpurun_out/bench_${TAG:-r01}_$n.json 2> gpurun_out/bench_${TAG:-r01}_$n.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG:-r01}_$n.json"))
print($n, "cold ms/step", d["ms_per_step"], "chained", d["chained"]["ms_per_step"], d["roofline"]["per_kernel_ms"], d["roofline"]["step_contract"])
PY
done
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct"
for k in row_kernel str_kernel; do ncu $M --clock-control none -k regex:^$k -s 4 -c 1 python bench.py --grid 8192 --steps 6 --warmup 3 --no-cpu --no-extra 2>&1 | grep -E "dram__|duration|hit_rate|void"; done
for k in row_kernel str_kernel; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$k -s 8 -c 1 -f -o gpurun_out/prof_${k}_${TAG:-r01}_1024 python bench.py --grid 1024 --steps 12 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_${k}_1024.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^str_kernel -s 4 -c 1 -f -o gpurun_out/prof_str_kernel_${TAG:-r01}_8192 python bench.py --grid 8192 --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_full_str_8192.log 2>&1
du -sh gpurun_out
