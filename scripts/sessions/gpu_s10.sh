#!/bin/bash
# Hardware check of the float build (libmifgpu_f32.so): the parity cases of tests/fp32_cases.py and the float host layer.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
timeout 75 python -m pytest tests/test_gpu_zzz_fp32.py -m gpu -q -x > gpurun_out/r02s10_pytest_fp32.log 2>&1; tail -15 gpurun_out/r02s10_pytest_fp32.log | cut -c1-300
MIFGPU_LIB=libmifgpu_f32.so timeout 30 python tests/fp32_cases.py > gpurun_out/r02s10_fp32_cases.json 2> gpurun_out/r02s10_fp32_cases.err; cut -c1-1500 gpurun_out/r02s10_fp32_cases.json
