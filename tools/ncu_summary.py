"""Print the handful of ncu metrics we track from a .ncu-rep (raw page), plus the warp-stall breakdown.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [launch_index]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__icc_request_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
]


def main():
    rep = sys.argv[1]
    idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, row = rows[0], rows[1], rows[2 + idx]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"kernel: {row[col['Kernel Name']]}")
    for k in KEYS:
        if k in col:
            print(f"{k:75s} {row[col[k]]:>16s} {units[col[k]]}")
    stalls = []
    for h, i in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(row[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    print("warp stalls (warps per issue-active cycle):")
    for v, n in sorted(stalls, reverse=True)[:10]:
        print(f"    {n:35s} {v:8.3f}")


if __name__ == "__main__":
    main()
