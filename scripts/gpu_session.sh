#!/bin/bash
# One gpurun call that brings back everything a tuning session needs (1 GPU):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_session.sh [tag]'
# 1. the GPU test suite, 2. the bench line (full step) and the Poisson-only line, 3. the ncu launch list of two steps,
# 4. one `ncu --set full` capture of the kernels of one step (source view on), 5. A/B lines for the switches named in
# $MIF_AB (space separated NAME=VALUE pairs; default: MIFGPU_NO_PLAIN_STRIDED=1, the pre-r01-final strided instantiation).
# Everything lands in gpurun_out/<tag>_*; scripts/summarize_ncu.py turns the captures into profiles/ files here.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
echo "== pytest -m gpu" && timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
echo "== bench (full step)" && timeout 600 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; tail -c 1500 $out/${tag}_bench_1gpu.json
echo "== bench (Poisson only)" && timeout 300 python bench.py --workload poisson --steps 10 --warmup 3 > $out/${tag}_bench_poisson.json 2>> $out/${tag}_bench_1gpu.err
for ab in ${MIF_AB:-MIFGPU_NO_PLAIN_STRIDED=1 MIFGPU_X_MIRROR_SHFL=1}; do
  echo "== A/B $ab" && env "$ab" timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > "$out/${tag}_bench_${ab//[^A-Za-z0-9_=]/_}.json" 2>> $out/${tag}_bench_1gpu.err
done
# 27 kernels per step, 3 warm-up steps: skip 81 launches, list two steps
echo "== ncu launch list" && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 81 -c 54 --csv \
  --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1
echo "== ncu --set full (one step)" && timeout 900 ncu --set full --clock-control none --import-source on -s 81 -c 27 \
  -o $out/${tag}_step python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1
# the SIMT interpreter cannot see races: racecheck the kernels that were written without a GPU (M = 2048 warp kernel,
# two-stage generic passes, the A/B variants) on thin grids
echo "== compute-sanitizer racecheck" && timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis \
  python -m pytest tests/test_gpu_zz_new_sizes.py -m gpu -q -x -k "test_pressure_solve_random_velocity_new_sizes" \
  > $out/${tag}_racecheck.log 2>&1; tail -5 $out/${tag}_racecheck.log
echo "== racecheck (strided sweeps)" && timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis \
  python -m pytest tests/test_gpu_vs_oracle.py -m gpu -q -x -k "test_pressure_solve_random_velocity and (N15 or N16)" \
  > $out/${tag}_racecheck_plain.log 2>&1; tail -3 $out/${tag}_racecheck_plain.log
ls -la $out | tail -20
