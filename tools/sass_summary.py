"""Per-kernel SASS evidence of libtsvpp.so (no GPU needed): registers / shared memory / spills from
`cuobjdump -res-usage`, and counts of the mnemonics that show what the kernels are built from
(UBLKCP = cp.async.bulk through the TMA unit, SYNCS = mbarrier, IDP = dp4a, I2IP / F2IP = packed saturating
conversions, DFMA = the FP64 tie path, LDS/STG...).  Usage: python tools/sass_summary.py [lib] > profiles/sass_summary.txt"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "tensor_stream_b200/libtsvpp.so"
WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "IDP", "I2IP", "F2IP", "I2FP", "PRMT", "DFMA", "DMUL", "LDS", "STS", "LDG", "STG", "ST", "BAR", "HMMA", "UTC"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for ln in res.splitlines():
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in ln:
            usage[cur] = ln.strip()
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w in ("UTC",) and op.startswith(w)):
                    counts[cur][w] += 1
    names = demangle(list(counts))
    print("# SASS summary of %s (sm_100a)" % LIB)
    print("# kernel | instructions | " + " ".join(WATCH) + " | resources")
    tot = collections.Counter()
    for k, c in counts.items():
        short = re.sub(r"\(.*", "", names.get(k, k)).replace("void tsvpp::", "")
        print("%-58s %6d | %s | %s" % (short[:58], c["_total"], " ".join("%s=%d" % (w, c[w]) for w in WATCH if c[w]), usage.get(k, "")))
        tot.update(c)
    print("# all kernels: %d functions, %d SASS instructions; %s" % (len(counts), tot["_total"], ", ".join("%s %d" % (w, tot[w]) for w in WATCH if tot[w])))
    print("# tensor-core mnemonics (HMMA / UTC*MMA): %d -- by design: the path is byte-granular resampling, not a contraction" % (tot["HMMA"] + tot["UTC"]))


if __name__ == "__main__":
    main()
