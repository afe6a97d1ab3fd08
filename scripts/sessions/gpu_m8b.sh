#!/bin/bash
# 8 GPUs, final: the pencil case that failed before the staging-buffer fix, all 8-GPU parity tests, the weak bench line.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
out=gpurun_out; tag=r02m8b; mkdir -p $out
echo "== multi-GPU parity tests that need 8 GPUs"
MIFGPU_REQUIRE_TMA=1 timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_zz_multi_pencils_drivers.py -m gpu -q -rs -k "8-2 or 8-4 or 2-4" > $out/${tag}_pytest_multi.log 2>&1; tail -6 $out/${tag}_pytest_multi.log | cut -c1-200
echo "== bench weak N=8"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 10 --warmup 3 > $out/${tag}_bench_weak.json 2> $out/${tag}_bench_weak.err
python - <<PY
import json
d=json.loads([l for l in open("$out/${tag}_bench_weak.json") if l.startswith("{")][-1]); print("   ms/step", round(d["ms_per_step"],3), "value", d["value"], "parity", d["parity_vs_single_rank"]["max_rel_linf"], "e2e", d["e2e"]["value"], "nvlink", d["nvlink"]["GBs_over_carrier_kernels"]); print("   ", d["kernels"])
PY
