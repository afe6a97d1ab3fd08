#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_multi.py -m gpu -q -x ) > gpurun_out/pytest_mg2.log 2>&1
tail -6 gpurun_out/pytest_mg2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_mg2_r2.json 2> gpurun_out/bench_mg2_r2.err
tail -c 2500 gpurun_out/bench_mg2_r2.json
