"""Extract the metrics the roofline numbers are derived from out of an .ncu-rep (run where ncu is installed, no GPU
needed):  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.csv"""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    keys = [k for k in KEYS if k in idx]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch"] + keys)
        w.writerow(["unit"] + [units[idx[k]] for k in keys])
        for n, r in enumerate(data):
            w.writerow([n] + [r[idx[k]] for k in keys])
    print("wrote", out, len(data), "launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
