#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_fdiv.log 2>&1
tail -4 gpurun_out/pytest_fdiv.log
show() {
python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
k=d["kernels"]
print(sys.argv[1].split("/")[-1], "ms/step", round(d["ms_per_step"],3), {n:k[n] for n in k if n.startswith("sweep") or n.startswith("div")})
PY
}
B="python bench.py --steps 5 --no-cpu-baseline --no-e2e"
$B > gpurun_out/fd_A.json 2> gpurun_out/fd_A.err; show gpurun_out/fd_A.json
MIFGPU_FUSED_DIVERGENCE=0 $B > gpurun_out/fd_B.json 2> gpurun_out/fd_B.err; show gpurun_out/fd_B.json
$B > gpurun_out/fd_A2.json 2> gpurun_out/fd_A2.err; show gpurun_out/fd_A2.json
