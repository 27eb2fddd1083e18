"""One gpurun call: [headline bench], ncu launch list, ncu --set full of the hot kernels.
Usage on the GPU box (repo root):  python ncu_capture.py <tag> [--no-bench] [--kernels a,b] [bench args...]"""
import os
import subprocess
import sys

args = sys.argv[1:]
tag = args[0] if args else "r01"
kernels = ["row_kernel", "str_kernel", "str_tma_kernel"]
extra = []
i = 1
nobench = False
while i < len(args):
    if args[i] == "--no-bench":
        nobench = True
    elif args[i] == "--kernels":
        kernels = args[i + 1].split(",")
        i += 1
    else:
        extra.append(args[i])
    i += 1
os.makedirs("gpurun_out", exist_ok=True)


def sh(cmd, timeout=None):
    print("+", cmd, flush=True)
    try:
        return subprocess.run(cmd, shell=True, timeout=timeout).returncode
    except subprocess.TimeoutExpired:
        print("TIMEOUT", cmd, flush=True)
        return -1


if not nobench:
    sh(f"python bench.py --steps 1000 --warmup 10 {' '.join(extra)} > gpurun_out/bench_{tag}.json 2> gpurun_out/bench_{tag}.err", 900)
    sh(f"tail -c 7000 gpurun_out/bench_{tag}.json; tail -5 gpurun_out/bench_{tag}.err")
small = "python bench.py --steps 30 --warmup 3 --no-cpu --no-extra " + " ".join(extra)
sh(f"ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 120 --csv "
   f"--log-file gpurun_out/launches_{tag}.csv {small} > gpurun_out/ncu_launch_{tag}.log 2>&1", 600)
for k in kernels:
    sh(f"ncu --set full --clock-control none --import-source on -k regex:^{k} -s 8 -c 1 -f "
       f"-o gpurun_out/prof_{k}_{tag} {small.replace('--steps 30', '--steps 12')} > gpurun_out/ncu_full_{k}_{tag}.log 2>&1", 900)
sh("ls -la gpurun_out")
