#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests/test_gpu_vs_oracle.py tests/test_gpu_golden.py -m gpu -q -x > gpurun_out/pytest_x.log 2>&1
tail -3 gpurun_out/pytest_x.log
for r in 1 2; do
python bench.py --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/x_$r.json 2> gpurun_out/x_$r.err
python - $r <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/x_{sys.argv[1]}.json"))
print("ms/step", round(d["ms_per_step"],3), d["kernels"])
PY
done
