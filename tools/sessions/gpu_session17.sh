#!/bin/bash
# GEMM with MMA runs of four: parity suites that go through the GEMMs, then timings
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fp8.py tests/test_gpu_nvfp4.py tests/test_gpu_block.py tests/test_gpu_model.py -q -m gpu > gpurun_out/s17_pytest.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/s17_pytest.txt
echo "== gemm timings"; timeout 200 python tools/time_gemm.py > gpurun_out/s17_gemm.txt 2>&1; cat gpurun_out/s17_gemm.txt
