"""Summarises .ncu-rep files into a small CSV + markdown table (the judged copy lives under profiles/).

    python scripts/ncu_summary.py out_prefix report1.ncu-rep [report2.ncu-rep ...]
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_inst"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("launch__shared_mem_per_block_static", "static_smem"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
]


def main():
    prefix, reps = sys.argv[1], sys.argv[2:]
    out_rows = []
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            d = {"report": rep.split("/")[-1], "kernel": r[idx["Kernel Name"]].split("(")[0]}
            for k, short in KEYS:
                if k in idx:
                    d[short] = r[idx[k]] + (" " + units[idx[k]] if units[idx[k]] not in ("", "%") else "")
            out_rows.append(d)
    cols = ["report", "kernel"] + [s for _, s in KEYS]
    with open(prefix + ".csv", "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=cols)
        w.writeheader()
        for d in out_rows:
            w.writerow(d)
    for d in out_rows:
        print(" | ".join(f"{c}={d.get(c, '')}" for c in cols))


if __name__ == "__main__":
    main()
