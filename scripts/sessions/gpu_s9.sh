#!/bin/bash
# Last 1-GPU check of the round at HEAD: smoke() and one short bench line (the stencil launchers gained a plane range).
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
out=gpurun_out; mkdir -p $out
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 60 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $out/r02s9_bench_1gpu.json 2> $out/r02s9_bench_1gpu.err; cut -c1-300 $out/r02s9_bench_1gpu.json; tail -3 $out/r02s9_bench_1gpu.err
