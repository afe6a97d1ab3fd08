#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_split3.log 2>&1
tail -5 gpurun_out/pytest_split3.log
MIFGPU_SPLIT_LINES_X=8 MIFGPU_SPLIT_LINES_YZ=8 python -m pytest tests/test_gpu_vs_oracle.py -m gpu -q -k "pressure_solve" > gpurun_out/pytest_split3b.log 2>&1
tail -3 gpurun_out/pytest_split3b.log
show() {
python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
k=d["kernels"]
print(sys.argv[1].split("/")[-1], "ms/step", round(d["ms_per_step"],3), {n:k[n] for n in k if n.startswith("sweep") or n.startswith("correct")})
PY
}
B="python bench.py --steps 4 --no-cpu-baseline --no-e2e"
$B --dims 1025 1025 129 > gpurun_out/s3_a44.json 2> gpurun_out/s3_a44.err; show gpurun_out/s3_a44.json
MIFGPU_SPLIT_LINES_X=8 MIFGPU_SPLIT_LINES_YZ=8 $B --dims 1025 1025 129 > gpurun_out/s3_a88.json 2> gpurun_out/s3_a88.err; show gpurun_out/s3_a88.json
$B --dims 129 1025 1025 > gpurun_out/s3_b44.json 2> gpurun_out/s3_b44.err; show gpurun_out/s3_b44.json
MIFGPU_SPLIT_LINES_X=8 MIFGPU_SPLIT_LINES_YZ=8 $B --dims 129 1025 1025 > gpurun_out/s3_b88.json 2> gpurun_out/s3_b88.err; show gpurun_out/s3_b88.json
$B > gpurun_out/s3_513.json 2> gpurun_out/s3_513.err; show gpurun_out/s3_513.json
