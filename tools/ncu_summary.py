"""Condense an ncu report (.ncu-rep, read with `ncu -i ... --page raw --csv`) into the handful of
numbers the roofline discussion needs.  Usage: python tools/ncu_summary.py REP [OUT.md] [frames]"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else None
    frames = int(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append("### %s  grid=%s block=%s" % (d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
        lines.append("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d:
                lines.append("| %s | %s | %s |" % (k, d[k], u[k]))
        try:
            rd = float(d["dram__bytes_read.sum"]) * (1e6 if u["dram__bytes_read.sum"] == "Mbyte" else 1e3 if u["dram__bytes_read.sum"] == "Kbyte" else 1e9 if u["dram__bytes_read.sum"] == "Gbyte" else 1)
            wr = float(d["dram__bytes_write.sum"]) * (1e6 if u["dram__bytes_write.sum"] == "Mbyte" else 1e3 if u["dram__bytes_write.sum"] == "Kbyte" else 1e9 if u["dram__bytes_write.sum"] == "Gbyte" else 1)
            lines.append("\ndram bytes per launch: read %.0f + write %.0f = %.0f" % (rd, wr, rd + wr))
            if frames:
                lines.append("per frame (%d frames/launch): read %.0f + write %.0f = %.0f B" % (frames, rd / frames, wr / frames, (rd + wr) / frames))
                lines.append("JSON: " + json.dumps({"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "frames_per_launch": frames}))
        except (KeyError, ValueError):
            pass
        lines.append("")
    txt = "\n".join(lines)
    if out:
        open(out, "w").write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    main()
