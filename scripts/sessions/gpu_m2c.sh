#!/bin/bash
# 2 GPUs: NCCL point-to-point channel count against the halo-exchange time (9 plane exchanges per step).
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
out=gpurun_out; tag=r02m2c
mkdir -p $out
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("$out/${tag}_$name.json") if l.startswith("{")][-1]); print("   $name ms/step", round(d["ms_per_step"],3), "halo", d["kernels"].get("halo_exchange"), "y_fwd", d["kernels"].get("sweep_y_fwd"), "z", d["kernels"].get("sweep_z_fused"))
except Exception as e: print("   no line:", e); print(open("$out/${tag}_$name.err").read()[-800:])
PY
}
run default --steps 10 --warmup 3 --no-e2e
NCCL_MIN_P2P_NCHANNELS=8 run min8 --steps 10 --warmup 3 --no-e2e
NCCL_MIN_P2P_NCHANNELS=16 run min16 --steps 10 --warmup 3 --no-e2e
NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 run min32 --steps 10 --warmup 3 --no-e2e
NCCL_MIN_P2P_NCHANNELS=16 NCCL_BUFFSIZE=16777216 run min16_buf16 --steps 10 --warmup 3 --no-e2e
