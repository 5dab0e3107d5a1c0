#!/usr/bin/env python
"""Turn an ncu report (--set full) into the compact per-kernel summary committed under profiles/.
Usage: python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/rNN_tag_ncu_summary.md "command line that was profiled"
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "smsp__inst_executed.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "sm__cycles_elapsed.max",
]


def main():
    rep, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none summary\n\nsource report: `{rep}` (scratch, not committed)\n\ncommand: `{cmd}`\n\n")
        f.write("Per-launch values; times under ncu are cold-cache and serialised (compare shares, not absolutes).\n")
        for r in data:
            name = r[hdr.index("Kernel Name")]
            f.write(f"\n## {name}\n\n| metric | value | unit |\n|---|---|---|\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"| {w} | {r[i]} | {units[i]} |\n")
            try:
                rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", ""))
                wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
                u = units[hdr.index("dram__bytes_read.sum")]
                f.write(f"| **traffic = dram read + write** | {rd + wr:.6f} | {u} |\n")
            except Exception:
                pass
    print("wrote", out)


if __name__ == "__main__":
    main()
