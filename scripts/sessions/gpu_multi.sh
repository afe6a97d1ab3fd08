#!/bin/bash
# Multi-GPU session on N GPUs: parity tests selected by $3 (pytest -k), weak / strong 1025^3 / aniso bench lines.
#   gpurun --gpus N -- 'bash scripts/sessions/gpu_multi.sh N tag "<pytest -k expression>"'
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
N=${1:-4}; tag=${2:-r02m$N}; sel=${3:-}
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/${tag}_gpus.txt
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("$out/${tag}_$name.json") if l.startswith("{")][-1]); print("   ms/step", round(d["ms_per_step"],3), "value", d["value"], "parity", d.get("parity_vs_single_rank",{}).get("max_rel_linf"), "e2e", d["e2e"] and d["e2e"].get("value"), "nvlink", d["nvlink"]["GBs_over_carrier_kernels"]); print("   ", d["kernels"])
except Exception as e: print("   no line:", e); print(open("$out/${tag}_$name.err").read()[-1500:])
PY
}
echo "== bench weak N=$N"; run bench_weak --steps 10 --warmup 3
echo "== bench strong 1025^3 N=$N"; run bench_strong1025 --scaling strong --size 1025 --steps 8 --warmup 3 --no-e2e
echo "== bench aniso N=$N"; run bench_aniso --workload aniso --steps 10 --warmup 3 --no-e2e
if [ -n "$sel" ]; then
  echo "== multi-GPU parity tests: $sel"
  MIFGPU_REQUIRE_TMA=1 timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_zz_multi_pencils_drivers.py -m gpu -q -rs -k "$sel" > $out/${tag}_pytest_multi.log 2>&1; tail -15 $out/${tag}_pytest_multi.log | cut -c1-200
fi
ls -la $out | tail -8
