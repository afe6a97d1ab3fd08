#!/bin/bash
# GPU experiment: parity + timing of the z-marching RK stage kernel against the plain one.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
MIFGPU_STAGE_MARCH=16 python -m pytest tests/test_gpu_vs_oracle.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/pytest_march.log 2>&1
tail -3 gpurun_out/pytest_march.log
for m in 0 4 16 64 600; do
  MIFGPU_STAGE_MARCH=$m python bench.py --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/march_$m.json 2> gpurun_out/march_$m.err
  python - <<PY
import json
d=json.load(open("gpurun_out/march_$m.json"))
k=d["kernels"]
print("march=$m", round(d["ms_per_step"],3), k["stage1"], k["stage2"], k["stage3"])
PY
done
