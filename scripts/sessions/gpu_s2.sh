#!/bin/bash
# Round-2 session 2 (1 GPU): full GPU suite with the TMA sweeps (64-byte swizzle), bench lines of configs 2/3/4(N=1)/5(N=1),
# pipelined e2e, reference arm on the 513^3 grid.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
tag=${1:-r02s2}
out=gpurun_out
mkdir -p $out
echo "== pytest -m gpu (TMA required where it applies)"
MIFGPU_REQUIRE_TMA=1 timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -4 $out/${tag}_pytest.log
echo "== bench (full step, default)"
timeout 900 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; tail -c 2500 $out/${tag}_bench_1gpu.json; tail -3 $out/${tag}_bench_1gpu.err
echo "== bench (Poisson only, config 2)"
timeout 300 python bench.py --workload poisson --steps 10 --warmup 3 > $out/${tag}_bench_poisson.json 2>> $out/${tag}_bench_1gpu.err; cut -c1-900 $out/${tag}_bench_poisson.json
echo "== bench (aniso N=1, config 5)"
timeout 300 python bench.py --workload aniso --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_aniso_1gpu.json 2>> $out/${tag}_bench_1gpu.err; cut -c1-500 $out/${tag}_bench_aniso_1gpu.json
echo "== bench (1025^3 on one GPU, config 4 N=1)"
timeout 900 python bench.py --scaling strong --size 1025 --steps 5 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_1025_1gpu.json 2>> $out/${tag}_bench_1gpu.err; cut -c1-700 $out/${tag}_bench_1025_1gpu.json; tail -2 $out/${tag}_bench_1gpu.err
echo "== reference arm (host CPU)"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench_1gpu.err; cut -c1-900 $out/${tag}_bench_reference.json
nproc; free -g | head -2
ls -la $out | tail -12
