#!/bin/bash
# Session 8 (1 GPU): x sweeps through per-warp TMA pipelines against the register-fed x sweeps (A/B), parity of the solve tests.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
out=gpurun_out; mkdir -p $out; : > $out/r02s8_ab.jsonl
for ab in "X=0" "MIFGPU_X_NO_TMA=1" "X=1" "MIFGPU_X_NO_TMA=1"; do
  env "$ab" timeout 300 python scripts/ab_timing.py 513 10 "$ab" >> $out/r02s8_ab.jsonl 2>> $out/r02s8.err
  tail -1 $out/r02s8_ab.jsonl | cut -c1-420
done
MIFGPU_REQUIRE_TMA=1 timeout 900 python -m pytest tests/test_gpu_vs_oracle.py tests/test_gpu_golden.py tests/test_gpu_zz_full_size.py -m gpu -x -q > $out/r02s8_pytest.log 2>&1; tail -3 $out/r02s8_pytest.log
