#!/bin/bash
# Round-2 session 3 (1 GPU): the 16 x 32 transform (mif_fft512.cuh) against the radix-8 passes, parity, ncu.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
tag=${1:-r02s3}
out=gpurun_out
mkdir -p $out
echo "== parity"
MIFGPU_REQUIRE_TMA=1 timeout 900 python -m pytest tests/test_gpu_vs_oracle.py tests/test_gpu_golden.py tests/test_gpu_zz_full_size.py -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -4 $out/${tag}_pytest.log
: > $out/${tag}_ab.jsonl
for ab in "X=0" "MIFGPU_FFT_RADIX8=1" "X=1" "MIFGPU_TMA_CTAS_PER_SM=1"; do
  echo "== A/B $ab"
  env "$ab" timeout 300 python scripts/ab_timing.py 513 10 "$ab" >> $out/${tag}_ab.jsonl 2>> $out/${tag}_ab.err
  tail -1 $out/${tag}_ab.jsonl | cut -c1-600
done
echo "== ncu --set full (sweep kernels of one solve)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dct512" -s 15 -c 5 -o $out/${tag}_sweeps \
  python scripts/ab_timing.py 513 1 ncu > $out/${tag}_ncu_full.log 2>&1
tail -2 $out/${tag}_ncu_full.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err; cut -c1-300 $out/${tag}_bench_1gpu.json; tail -2 $out/${tag}_bench_1gpu.err
ls -la $out | tail -8
