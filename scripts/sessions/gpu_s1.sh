#!/bin/bash
# Round-2 session 1 (1 GPU): first hardware run of the TMA-staged strided sweeps -- parity, A/B timings, ncu.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
tag=${1:-r02s1}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt
echo "== parity (TMA required)"
MIFGPU_REQUIRE_TMA=1 MIFGPU_TMA_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_vs_oracle.py tests/test_gpu_golden.py -m gpu -x -q > $out/${tag}_pytest_tma.log 2>&1; tail -5 $out/${tag}_pytest_tma.log
echo "== parity (plain output map)"
MIFGPU_TMA_NO_SWIZZLE=1 MIFGPU_REQUIRE_TMA=1 timeout 600 python -m pytest tests/test_gpu_vs_oracle.py -m gpu -x -q -k "test_pressure_solve_random_velocity" > $out/${tag}_pytest_noswz.log 2>&1; tail -3 $out/${tag}_pytest_noswz.log
: > $out/${tag}_ab.jsonl
for ab in "X=0" "MIFGPU_NO_TMA=1" "MIFGPU_TMA_NO_SWIZZLE=1" "MIFGPU_TMA_L2PROMO=0" "MIFGPU_TMA_L2PROMO=3" "MIFGPU_TMA_CTAS_PER_SM=1" "MIFGPU_X_MIRROR_SHFL=1" "X=1"; do
  echo "== A/B $ab"
  env "$ab" timeout 300 python scripts/ab_timing.py 513 10 "$ab" >> $out/${tag}_ab.jsonl 2>> $out/${tag}_ab.err
  tail -1 $out/${tag}_ab.jsonl | cut -c1-600
done
echo "== ncu launch list (one step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 81 -c 27 --csv --log-file $out/${tag}_launches.csv \
  python scripts/ab_timing.py 513 1 ncu > $out/${tag}_ncu_launches.log 2>&1
echo "== ncu --set full (TMA sweeps)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tma_dct -s 9 -c 3 -o $out/${tag}_tma \
  python scripts/ab_timing.py 513 1 ncu > $out/${tag}_ncu_full.log 2>&1
tail -3 $out/${tag}_ncu_full.log
echo "== memcheck (thin grids)"
MIFGPU_REQUIRE_TMA=1 timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_vs_oracle.py -m gpu -q -x -k "test_pressure_solve_random_velocity and (N15 or N16)" > $out/${tag}_memcheck.log 2>&1; tail -4 $out/${tag}_memcheck.log
ls -la $out | tail -15
