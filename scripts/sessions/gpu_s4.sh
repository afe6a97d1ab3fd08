#!/bin/bash
# Round-2 session 4 (1 GPU): staggered start of the second CTA per SM in the TMA sweeps (A/B), full GPU suite.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
tag=${1:-r02s4}
out=gpurun_out
mkdir -p $out
: > $out/${tag}_ab.jsonl
for ab in "X=0" "MIFGPU_TMA_STAGGER_NS=1000" "MIFGPU_TMA_STAGGER_NS=2000" "MIFGPU_TMA_STAGGER_NS=3000" "MIFGPU_TMA_STAGGER_NS=5000" "X=1"; do
  echo "== A/B $ab"
  env "$ab" timeout 300 python scripts/ab_timing.py 513 10 "$ab" >> $out/${tag}_ab.jsonl 2>> $out/${tag}_ab.err
  tail -1 $out/${tag}_ab.jsonl | cut -c1-560
done
echo "== pytest -m gpu"
MIFGPU_REQUIRE_TMA=1 timeout 1800 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -4 $out/${tag}_pytest.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
