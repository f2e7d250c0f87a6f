"""Turn gpurun_out/*.ncu-rep into the per-kernel text summaries committed under profiles/ (run in the build container:
`python profiles/summarize.py gpurun_out/prof_x.ncu-rep profiles/r1_x.summary.csv`)."""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(hdr) if h in WANT or h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows) - 2)])
        for c in cols:
            w.writerow([hdr[c], units[c]] + [r[c] for r in rows[2:]])
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
