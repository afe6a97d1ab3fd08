#!/bin/bash
# Round-2 session 5 (1 GPU): e2e pipeline depth A/B, ncu --set full of every kernel of one step at HEAD, final bench lines.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
tag=${1:-r02s5}
out=gpurun_out
mkdir -p $out
for sets in 2 3 4; do
  echo "== e2e with $sets device field sets"
  MIF_BENCH_E2E_SETS=$sets timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $out/${tag}_e2e_sets$sets.json 2>> $out/${tag}.err
  python - <<PY
import json
d=json.load(open("$out/${tag}_e2e_sets$sets.json")); print("   step ms", round(d["ms_per_step"],3), "e2e", {k:v for k,v in d["e2e"].items() if k!="what"})
PY
done
echo "== ncu launch list of one step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 81 -c 27 --csv --log-file $out/${tag}_launches.csv \
  python scripts/ab_timing.py 513 1 ncu > $out/${tag}_ncu_launches.log 2>&1
echo "== ncu --set full of one step"
timeout 1200 ncu --set full --clock-control none --import-source on -s 81 -c 27 -o $out/${tag}_step \
  python scripts/ab_timing.py 513 1 ncu > $out/${tag}_ncu_full.log 2>&1; tail -2 $out/${tag}_ncu_full.log
echo "== bench lines"
timeout 900 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_1gpu.json 2>> $out/${tag}.err; cut -c1-250 $out/${tag}_bench_1gpu.json
timeout 300 python bench.py --workload poisson --steps 10 --warmup 3 > $out/${tag}_bench_poisson.json 2>> $out/${tag}.err; cut -c1-250 $out/${tag}_bench_poisson.json
timeout 300 python bench.py --workload aniso --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_bench_aniso_1gpu.json 2>> $out/${tag}.err; cut -c1-250 $out/${tag}_bench_aniso_1gpu.json
timeout 900 python bench.py --scaling strong --size 1025 --steps 5 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_1025_1gpu.json 2>> $out/${tag}.err; cut -c1-250 $out/${tag}_bench_1025_1gpu.json
tail -3 $out/${tag}.err
ls -la $out | tail -12
