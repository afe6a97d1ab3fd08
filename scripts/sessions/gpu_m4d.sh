#!/bin/bash
# 4 GPUs: overlapped halo exchanges (second stream, interior planes first) against MIFGPU_NO_HALO_OVERLAP=1, with the
# bench's own multi-rank parity key
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
out=gpurun_out; tag=r02m4d
mkdir -p $out
run() { name=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("$out/${tag}_$name.json") if l.startswith("{")][-1]); print("   $name ms/step", round(d["ms_per_step"],3), "parity", d.get("parity_vs_single_rank",{}).get("max_rel_linf"), "halo", d["kernels"].get("halo_exchange")); print("   ", d["kernels"])
except Exception as e: print("   no line:", e); print(open("$out/${tag}_$name.err").read()[-1500:])
PY
}
run overlap --steps 30 --warmup 3 --no-e2e
MIFGPU_NO_HALO_OVERLAP=1 run no_overlap --steps 30 --warmup 3 --no-e2e
