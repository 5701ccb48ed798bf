"""Summarise an .ncu-rep (read here with `ncu -i`, no GPU needed) into a small JSON under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_summary_rNN.json
"""
import csv
import io
import json
import subprocess
import sys

WANT = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    summary = [{w: (r[i] + (" " + units[i] if units[i] else "")).strip() for w, i in idx}
               for r in rows[2:]]
    json.dump(summary, open(out, "w"), indent=1)
    for k in summary:
        print(k["Kernel Name"][:60], k.get("gpu__time_duration.sum"),
              "fma%", k.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
              "dram r/w", k.get("dram__bytes_read.sum"), k.get("dram__bytes_write.sum"))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
