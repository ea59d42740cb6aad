"""Developer sweep over the plan knobs (env overrides read by tsvpp_capi.cpp at plan build)."""
import itertools
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
combos = []
for rp, ns, tw in itertools.product((1, 2), (2, 3, 4, 6), (256, 128)):
    combos.append({"TSVPP_RP": rp, "TSVPP_STAGES": ns, "TSVPP_TW": tw})
extra = [{"TSVPP_RP": 2, "TSVPP_STAGES": 3, "TSVPP_CTAS_PER_SM": 2}, {"TSVPP_RP": 2, "TSVPP_STAGES": 3, "TSVPP_CTAS_PER_SM": 1},
         {"TSVPP_RP": 2, "TSVPP_STAGES": 3, "QB_FRAMES": 768}, {"TSVPP_RP": 1, "TSVPP_STAGES": 4, "QB_FRAMES": 768}]
for c in combos + extra:
    env = dict(os.environ)
    env.update({k: str(v) for k, v in c.items()})
    env.setdefault("QB_FRAMES", "192")
    env["QB_ITERS"] = "30"
    out = subprocess.run([sys.executable, os.path.join(HERE, "quick_bench.py")], env=env, capture_output=True, text=True)
    lines = [l for l in out.stdout.splitlines() if "fps" in l or "smem=" in l]
    plan = lines[0].split("|")[1].strip() if lines else out.stderr[-300:]
    keep = " ".join(x for x in plan.split() if x.split("=")[0] in ("tw", "rp", "nstages", "smem", "ctas_per_sm"))
    print(c, "|", keep, "|", lines[1].strip() if len(lines) > 1 else "FAILED", flush=True)
