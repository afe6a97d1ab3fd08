#!/bin/bash
# 2 GPUs: the TMA peer path on hardware -- parity tests that fit two devices, weak / strong / aniso bench lines, A/B against the LSU peer kernels.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
tag=${1:-r02m2b}
out=gpurun_out
mkdir -p $out
echo "== multi-GPU parity tests (world 2)"
MIFGPU_REQUIRE_TMA=1 timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_zz_multi_pencils_drivers.py -m gpu -q -rs -k "not pencil_decomposition" > $out/${tag}_pytest_multi.log 2>&1; tail -12 $out/${tag}_pytest_multi.log | cut -c1-200
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("$out/${tag}_$name.json") if l.startswith("{")][-1]); print("   ms/step", round(d["ms_per_step"],3), "parity", d.get("parity_vs_single_rank",{}).get("max_rel_linf"), "e2e", d["e2e"] and d["e2e"].get("value"), "nvlink", d["nvlink"]["GBs_over_carrier_kernels"]); print("   ", d["kernels"])
except Exception as e: print("   no line:", e); print(open("$out/${tag}_$name.err").read()[-1500:])
PY
}
echo "== bench weak N=2"; run bench_weak --steps 10 --warmup 3
echo "== bench weak N=2, LSU peer kernels"; MIFGPU_NO_TMA_PEER=1 run bench_weak_lsu --steps 10 --warmup 3 --no-e2e
echo "== bench strong 1025^3 N=2"; run bench_strong1025 --scaling strong --size 1025 --steps 5 --warmup 3
echo "== bench aniso N=2"; run bench_aniso --workload aniso --steps 10 --warmup 3 --no-e2e
ls -la $out | tail -8
