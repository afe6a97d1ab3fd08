#!/bin/bash
# GPU experiment: parity of the split M = 1024 sweeps + timing against the two-warp variant; stage kernel y chunks.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests/test_gpu_vs_oracle.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/pytest_split.log 2>&1
tail -5 gpurun_out/pytest_split.log
run() {  # name, env..., -- bench args
  name=$1; shift
  env "$@" > /dev/null 2>&1 || true
}
show() {
python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
k=d["kernels"]
print(sys.argv[1].split("/")[-1], "ms/step", round(d["ms_per_step"],3), {n:k[n] for n in k})
PY
}
B="python bench.py --steps 4 --no-cpu-baseline --no-e2e"
$B --dims 1025 1025 129 > gpurun_out/s_a_split.json 2> gpurun_out/s_a_split.err; show gpurun_out/s_a_split.json
MIFGPU_FFT_NO_SPLIT=1 MIFGPU_STAGE_CHUNK_MB=100000 $B --dims 1025 1025 129 > gpurun_out/s_a_old.json 2> gpurun_out/s_a_old.err; show gpurun_out/s_a_old.json
$B --dims 129 1025 1025 > gpurun_out/s_b_split.json 2> gpurun_out/s_b_split.err; show gpurun_out/s_b_split.json
MIFGPU_FFT_NO_SPLIT=1 MIFGPU_STAGE_CHUNK_MB=100000 $B --dims 129 1025 1025 > gpurun_out/s_b_old.json 2> gpurun_out/s_b_old.err; show gpurun_out/s_b_old.json
$B > gpurun_out/s_513.json 2> gpurun_out/s_513.err; show gpurun_out/s_513.json
MIFGPU_STAGE_CHUNK_MB=1.1 $B > gpurun_out/s_513_c1.json 2> gpurun_out/s_513_c1.err; show gpurun_out/s_513_c1.json
