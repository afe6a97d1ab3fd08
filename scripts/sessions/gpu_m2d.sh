#!/bin/bash
# 2 GPUs: overlapped halo exchanges (second stream, interior planes first) against MIFGPU_NO_HALO_OVERLAP=1, with the
# bench's own multi-rank parity key, plus the world-2 parity tests of the slab path (late-arriving halos on real NCCL).
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
out=gpurun_out; tag=r02m2d
mkdir -p $out
run() { name=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("$out/${tag}_$name.json") if l.startswith("{")][-1]); print("   $name ms/step", round(d["ms_per_step"],3), "parity", d.get("parity_vs_single_rank",{}).get("max_rel_linf"), "halo", d["kernels"].get("halo_exchange")); print("   ", d["kernels"])
except Exception as e: print("   no line:", e); print(open("$out/${tag}_$name.err").read()[-1500:])
PY
}
run overlap --steps 20 --warmup 3 --no-e2e
MIFGPU_NO_HALO_OVERLAP=1 run no_overlap --steps 20 --warmup 3 --no-e2e
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "2-full_16_2 or 2-lid1 or periodic_z and 2" > $out/${tag}_pytest.log 2>&1; tail -4 $out/${tag}_pytest.log | cut -c1-200
