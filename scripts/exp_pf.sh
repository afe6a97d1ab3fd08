#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_pf.log 2>&1
tail -4 gpurun_out/pytest_pf.log
show() {
python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
k=d["kernels"]
print(sys.argv[1].split("/")[-1], "ms/step", round(d["ms_per_step"],3), {n:k[n] for n in k if n.startswith("sweep") or n.startswith("stage")})
PY
}
B="python bench.py --steps 5 --no-cpu-baseline --no-e2e"
$B > gpurun_out/pf_A.json 2> gpurun_out/pf_A.err; show gpurun_out/pf_A.json
MIFGPU_STAGE_ONE_POINT=1 $B > gpurun_out/pf_B.json 2> gpurun_out/pf_B.err; show gpurun_out/pf_B.json
MIFGPU_SWEEP_PREFETCH=0 $B > gpurun_out/pf_C.json 2> gpurun_out/pf_C.err; show gpurun_out/pf_C.json
MIFGPU_SWEEP_PREFETCH=592 $B > gpurun_out/pf_D.json 2> gpurun_out/pf_D.err; show gpurun_out/pf_D.json
MIFGPU_SWEEP_PREFETCH=148 $B > gpurun_out/pf_E.json 2> gpurun_out/pf_E.err; show gpurun_out/pf_E.json
