"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion needs.
Usage: python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [substring filters...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "sm__inst_executed_pipe_",
    "smsp__average_warp", "smsp__average_warps_issue_stalled", "local", "sm__pipe_fma", "sm__pipe_fp64",
    "sm__inst_executed_pipe_lsu", "smsp__inst_executed_pipe", "lts__t_sector_hit_rate", "lts__t_bytes.sum ",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "achieved_occupancy", "sm__maximum_warps",
]


def main():
    rep = sys.argv[1]
    filt = sys.argv[2:]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    keys = filt or KEYS
    for r in data[:1]:
        print("====", r[hdr.index("Kernel Name")][:150])
        for i, h in enumerate(hdr):
            if any(k in h for k in keys):
                print(f"{h[:95]:95s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main()
