#!/bin/bash
# per-frame VideoProcessor::Convert rate through the C++ drop-in (C2 parameters, synthetic frame)
set -e
cd "$(dirname "$0")/.."
python - <<'PY'
import numpy as np
rng = np.random.default_rng(0)
rng.integers(0, 256, (1080 * 3 // 2, 1920), dtype=np.uint8).tofile("/tmp/c2_frame.nv12")
PY
for mode in 0 1; do
  echo "TSVPP_ASYNC=$mode"
  TSVPP_ASYNC=$mode tests/cpp/dropin_main /tmp/c2_frame.nv12 1920 1080 1920 2 0 1 256 256 1 0 0 0 0 20000 2>&1 | tail -4
done
