#!/bin/bash
# Multi-GPU companion of gpu_session.sh:   gpurun --gpus 8 --timeout 1500 -- 'bash scripts/gpu_session_multi.sh 8 [tag]'
# 1. the multi-GPU tests (slabs, pencils, ported drivers), 2. bench lines at N GPUs: default (peer-memory transposes),
# MIFGPU_SEG_CARRY=1 (segment cursors in the map lookups of the fused sweeps), MIFGPU_NO_PEER=1 (NCCL all-to-all), and
# --py 2 (Py x Pz pencils).  Never wrap these in ncu.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
N=${1:-8}
tag=${2:-r02}
out=gpurun_out
mkdir -p $out
run() {  # name, extra bench args, environment assignments...
  local name=$1 args=$2; shift 2
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29618 \
    bench.py --gpus $N --steps 5 --warmup 3 $args > $out/${tag}_bench_${N}gpu_${name}.json 2> $out/${tag}_bench_${N}gpu_${name}.err
  grep '^{' $out/${tag}_bench_${N}gpu_${name}.json | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('$name', d['ms_per_step'], d['value'], d.get('kernels'))"
}
echo "== pytest multi" && timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $out/${tag}_pytest_multi.log 2>&1; tail -3 $out/${tag}_pytest_multi.log
run default "" MIF_DUMMY=1
run segcarry "" MIFGPU_SEG_CARRY=1
run nopeer "" MIFGPU_NO_PEER=1
run pencils "--py 2" MIF_DUMMY=1
