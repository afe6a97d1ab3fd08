#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_vel.log 2>&1
tail -12 gpurun_out/pytest_vel.log
python bench.py --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/vel_A.json 2> gpurun_out/vel_A.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/vel_A.json"))
print("ms/step", round(d["ms_per_step"],3), d["kernels"])
PY
