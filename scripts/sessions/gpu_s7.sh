#!/bin/bash
# Session 7 (1 GPU): stage kernels compiled for 4 / 5 / 6 resident CTAs per SM (128 / 96 / 80 registers).
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
out=gpurun_out; mkdir -p $out; : > $out/r02s7_ab.jsonl
B=$PWD/mpi-incompressible-fluid_b200/build
for lib in "" $B/libmifgpu_stage5.so $B/libmifgpu_stage6.so "" $B/libmifgpu_stage5.so; do
  if [ -n "$lib" ]; then export MIFGPU_LIB=$lib; else unset MIFGPU_LIB; fi
  timeout 300 python scripts/ab_timing.py 513 10 "lib=${lib##*/}" >> $out/r02s7_ab.jsonl 2>> $out/r02s7.err
  tail -1 $out/r02s7_ab.jsonl | cut -c1-330
done
