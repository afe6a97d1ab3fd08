#!/bin/bash
# Round-2 multi-GPU session on 2 GPUs: the N > 1 parity tests that fit two devices, the 2-GPU bench lines (weak with the
# parity_vs_single_rank key, strong 1025^3, aniso), A/B of the segment-cursor lookups.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
tag=${1:-r02m2}
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/${tag}_gpus.txt
echo "== multi-GPU parity tests (world 2)"
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_zz_multi_pencils_drivers.py -m gpu -q -rs > $out/${tag}_pytest_multi.log 2>&1; tail -25 $out/${tag}_pytest_multi.log | cut -c1-200
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; cut -c1-400 $out/${tag}_$name.json; python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$name.json")); print("   ms/step", d["ms_per_step"], "parity", d.get("parity_vs_single_rank",{}).get("max_rel_linf"), "e2e", d["e2e"] and d["e2e"].get("value")); print("   ", d["kernels"])
except Exception as e: print("   no line:", e)
PY
}
echo "== bench weak N=2"; run bench_weak --steps 10 --warmup 3
echo "== bench weak N=2, segment cursors"; MIFGPU_SEG_CARRY=1 run bench_weak_segcarry --steps 10 --warmup 3 --no-e2e
echo "== bench strong 1025^3 N=2"; run bench_strong1025 --scaling strong --size 1025 --steps 5 --warmup 3
echo "== bench aniso N=2"; run bench_aniso --workload aniso --steps 10 --warmup 3
ls -la $out | tail -12
