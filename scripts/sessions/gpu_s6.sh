#!/bin/bash
# Round-2 session 6 (1 GPU): phase trace of the 513-point TMA sweep (diagnostic build), ncu --set full of the stencil kernels
# of one step, final bench lines (the reports of session 5 exceeded the 64 MiB that travel back).
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
tag=${1:-r02s6}
out=gpurun_out
mkdir -p $out
for mode in 0 2; do
  MIFGPU_LIB=$PWD/mpi-incompressible-fluid_b200/build/libmifgpu_trace.so MIFGPU_TRACE_MODE=$mode MIFGPU_TRACE_FILE=$out/${tag}_trace_mode$mode.bin \
    timeout 300 python scripts/ab_timing.py 513 3 trace$mode > $out/${tag}_trace_mode$mode.json 2>> $out/${tag}.err
  python scripts/phase_trace.py $out/${tag}_trace_mode$mode.bin 2>&1 | head -30
done
echo "== ncu --set full: stage / divergence / correct kernels of one step"
timeout 900 ncu --set full --clock-control none -k regex:"stage_kernel_pair|correct_kernel|divergence_kernel" -s 27 -c 9 -o $out/${tag}_stencils \
  python scripts/ab_timing.py 513 1 ncu > $out/${tag}_ncu_full.log 2>&1; tail -2 $out/${tag}_ncu_full.log
echo "== ncu launch list of one step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 81 -c 27 --csv --log-file $out/${tag}_launches.csv \
  python scripts/ab_timing.py 513 1 ncu > $out/${tag}_ncu_launches.log 2>&1
echo "== bench lines"
timeout 900 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_1gpu.json 2>> $out/${tag}.err; cut -c1-200 $out/${tag}_bench_1gpu.json
timeout 300 python bench.py --workload poisson --steps 10 --warmup 3 > $out/${tag}_bench_poisson.json 2>> $out/${tag}.err; cut -c1-200 $out/${tag}_bench_poisson.json
timeout 300 python bench.py --workload aniso --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_bench_aniso_1gpu.json 2>> $out/${tag}.err; cut -c1-200 $out/${tag}_bench_aniso_1gpu.json
timeout 900 python bench.py --scaling strong --size 1025 --steps 5 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_1025_1gpu.json 2>> $out/${tag}.err; cut -c1-200 $out/${tag}_bench_1025_1gpu.json
tail -3 $out/${tag}.err
du -sh $out; ls -la $out | tail -14
