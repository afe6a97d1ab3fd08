#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests/test_gpu_vs_oracle.py tests/test_gpu_golden.py -m gpu -q > gpurun_out/pytest_split2.log 2>&1
tail -15 gpurun_out/pytest_split2.log
MIFGPU_SPLIT_LINES_X=8 MIFGPU_SPLIT_LINES_YZ=8 python -m pytest tests/test_gpu_vs_oracle.py -m gpu -q -k "1025" > gpurun_out/pytest_split2b.log 2>&1
tail -3 gpurun_out/pytest_split2b.log
show() {
python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
k=d["kernels"]
print(sys.argv[1].split("/")[-1], "ms/step", round(d["ms_per_step"],3), {n:k[n] for n in k if n.startswith("sweep") or n.startswith("stage")})
PY
}
B="python bench.py --steps 4 --no-cpu-baseline --no-e2e"
$B --dims 1025 1025 129 > gpurun_out/s2_a44.json 2> gpurun_out/s2_a44.err; show gpurun_out/s2_a44.json
$B --dims 129 1025 1025 > gpurun_out/s2_b44.json 2> gpurun_out/s2_b44.err; show gpurun_out/s2_b44.json
