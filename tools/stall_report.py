"""Per-instruction view of an ncu capture exported on the GPU box by tools/ncu_cfg.sh
(X.raw.csv = `--page raw --csv`, X.src.csv = `--page source --csv --print-source sass`).
Usage: python tools/stall_report.py gpurun_out/X [frames] [top]"""
import csv
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    base = sys.argv[1]
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    rows = list(csv.reader(open(base + ".raw.csv")))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("kernel:", vals[hdr.index("Kernel Name")][:110])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("  %-70s %s %s" % (k, vals[i], units[i]))
    rows = list(csv.reader(open(base + ".src.csv")))
    h = rows[1]
    data = [r for r in rows[2:] if len(r) == len(h)]
    ia, ie = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    addr0 = int(data[0][0], 16)
    tot_s = sum(int(r[ia]) for r in data)
    tot_i = sum(int(r[ie]) for r in data)
    print("samples %d, warp instructions %d%s" % (tot_s, tot_i, (" (%.0f per frame)" % (tot_i / frames)) if frames else ""))
    cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    d = {c[6:]: sum(int(r[h.index(c)]) for r in data) for c in cols}
    print("stall reasons:", ", ".join("%s %d" % kv for kv in sorted(d.items(), key=lambda kv: -kv[1]) if kv[1] > tot_s * 0.01))
    ops = Counter()
    for r in data:
        t = r[1].split()
        op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
        ops[op.split(".")[0]] += int(r[ie])
    print("executed by opcode:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot_i) for k, v in ops.most_common(22)))
    print("top instructions by stall samples:")
    for r in sorted(data, key=lambda r: -int(r[ia]))[:top]:
        print("  %6s %-66s %6s %9s" % (hex(int(r[0], 16) - addr0), r[1].strip()[:66], r[ia], r[ie]))


if __name__ == "__main__":
    main()
