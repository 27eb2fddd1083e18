#!/bin/bash
# Round-2 evidence in one gpurun call (1 GPU): smoke, GPU test log, the default bench line (all extras), ncu launch list
# of the bench command, ncu --set full of the hot kernels of C2 (row / strided), C3 (component-parallel pair) and C4
# (Truncated-Wigner row kernel).   TAG=r02u bash tools/evidence_r02.sh
TAG=${TAG:-r02u}
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; tail -2 gpurun_out/smoke_${TAG}.log
( time timeout 900 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/pytest_gpu_${TAG}.log 2>&1
tail -12 gpurun_out/pytest_gpu_${TAG}.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 1500 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_${TAG}.err
small="python bench.py --steps 30 --warmup 3 --no-cpu --no-extra"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 120 --csv --log-file gpurun_out/launches_${TAG}.csv $small > gpurun_out/ncu_launch_${TAG}.log 2>&1
cap() {  # name kernel-regex bench-args...
  name=$1; k=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$k -s 8 -c 1 -f -o gpurun_out/prof_${name}_${TAG} python bench.py --steps 12 --warmup 3 --no-cpu --no-extra "$@" > gpurun_out/ncu_full_${name}_${TAG}.log 2>&1
  rep=gpurun_out/prof_${name}_${TAG}.ncu-rep
  { echo "# ncu --set full --clock-control none --import-source on -k regex:^$k -s 8 -c 1, inside: python bench.py --steps 12 --warmup 3 --no-cpu --no-extra $*"
    python tools/ncu_summary.py $rep; echo "## stall reasons (warp samples)"; python tools/ncu_stalls.py $rep
    echo "## hottest instructions"; python tools/ncu_hot.py $rep 20; echo "## executed instruction mix"; python tools/ncu_opmix.py $rep | head -30
  } > gpurun_out/ncu_${name}_${TAG}.txt 2>&1
  case $name in c2_str|c4_row) ;; *) rm -f $rep ;; esac   # gpurun brings back at most 64 MiB
}
cap c2_row row_kernel
cap c2_str str_kernel
cap c3_row row_cp_kernel --workload c3
cap c3_str str_cp_kernel --workload c3
cap c4_row row_kernel --workload c4 --nbatch 512
cap c4_str str_kernel --workload c4 --nbatch 512
ls -la gpurun_out | tail -30
du -sh gpurun_out
