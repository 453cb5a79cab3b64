#!/bin/bash
# 2-GPU re-check of the fused Ulysses / CFG-parallel / tile-parallel paths after the MMA-issue changes
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/test_gpu_ulysses.py -q -m gpu > gpurun_out/s18_pytest_2gpu.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/s18_pytest_2gpu.txt
