"""Markdown summary of `ncu --set full` reports: the metrics DESIGN.md argues with, per captured launch.

    python tools/ncu_summary.py title=path.ncu-rep [title=path.ncu-rep ...] > profiles/<round>_ncu_summary.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
    "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def main(args):
    print("# ncu summaries (B200, `ncu --set full --clock-control none --import-source on`; reports read with "
          "`ncu -i ... --page raw --csv`)\n")
    for arg in args:
        title, path = arg.rsplit("=", 1)
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"## {title}\n\n(no launch captured in {path})\n")
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            print(f"## {title}\n\n| metric | value |\n|---|---|")
            for k in KEYS:
                if k in d:
                    print(f"| `{k}` | {d[k]} {u.get(k, '')} |")
            print()


if __name__ == "__main__":
    main(sys.argv[1:])
