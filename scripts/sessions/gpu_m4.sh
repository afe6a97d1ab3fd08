#!/bin/bash
# 4 GPUs: multi-GPU parity tests that need four devices, the 4-GPU bench lines, and the Py = 4 pencil case that failed on 8.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
out=gpurun_out; tag=r02m4
mkdir -p $out
for py in 4 2; do
  echo "== pencils Py=$py on 4 ranks, lid1_12x10x14_2"
  MIF_PY=$py timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2965$py tests/mp_worker.py lid1_12x10x14_2 > $out/${tag}_pencil_py$py.log 2>&1
  grep -E "^\{|\[rank[0-9]\]:.*(Error|error|assert)|libmifgpu" $out/${tag}_pencil_py$py.log | head -8
done
bash scripts/sessions/gpu_multi.sh 4 r02m4 "4"
