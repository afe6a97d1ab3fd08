#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_mg${N}_r3.json 2> gpurun_out/bench_mg${N}_r3.err
tail -c 3000 gpurun_out/bench_mg${N}_r3.json
tail -5 gpurun_out/bench_mg${N}_r3.err
