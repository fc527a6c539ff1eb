"""Summarise an `ncu --set full` report: python scripts/ncu_summary.py gpurun_out/attn.ncu-rep "header line" > profiles/...txt
(reads the raw page through `ncu -i ... --page raw --csv`)."""
import csv
import subprocess
import sys

METRICS = [
    ("duration", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("cluster size", "launch__cluster_size"),
    ("tensor pipe active % (of active cycles)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor pipe active % (of elapsed)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("SM clock", "sm__cycles_elapsed.avg.per_second"),
    ("DRAM read", "dram__bytes_read.sum"),
    ("DRAM write", "dram__bytes_write.sum"),
    ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("issue active % per SMSP", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("IPC (all SMs)", "sm__inst_executed.sum.per_cycle_active"),
    ("regs/thread", "launch__registers_per_thread"),
    ("smem bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("stall long_scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall membar", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"),
    ("stall barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall math_pipe_throttle", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
]


def main(rep, header):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(header)
    print("(times under the profiler are cold-cache and serialised; the bench line's CUDA-event times are the ones reported)\n")
    for r in rows[2:]:
        print("== " + r[col["Kernel Name"]].split("(")[0])
        for label, m in METRICS:
            if m in col:
                print(f"   {label:42s} {r[col[m]]} {units[col[m]]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu --set full --clock-control none")
