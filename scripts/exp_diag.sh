#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_diag.log 2>&1
tail -6 gpurun_out/pytest_diag.log
show() {
python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
k=d["kernels"]
print(sys.argv[1].split("/")[-1], "ms/step", round(d["ms_per_step"],3), {n:k[n] for n in k})
PY
}
B="python bench.py --steps 5 --no-cpu-baseline --no-e2e"
$B > gpurun_out/dg_A.json 2> gpurun_out/dg_A.err; show gpurun_out/dg_A.json
python bench.py --workload poisson --steps 10 > gpurun_out/dg_poisson.json 2> gpurun_out/dg_poisson.err; cat gpurun_out/dg_poisson.json
MIFGPU_NO_WARP_RFFT=1 python bench.py --workload poisson --steps 10 > gpurun_out/dg_poisson_generic.json 2> gpurun_out/dg_poisson_generic.err; cat gpurun_out/dg_poisson_generic.json
