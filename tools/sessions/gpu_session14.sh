#!/bin/bash
# MMA runs of four under one election (FMHA + halo conv): parity first, then timings
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_vae.py -x -q -m gpu > gpurun_out/s14_tests.txt 2>&1
tail -4 gpurun_out/s14_tests.txt
{
timeout 100 python tools/time_fmha.py 75600 40 128 4
timeout 100 python tools/time_fmha.py 75600 5 128 8
timeout 100 python tools/time_fmha.py 17776 48 64 8
timeout 60 python tools/prof_conv.py 96 8 720 1280
timeout 60 python tools/prof_conv.py 192 8 720 1280 96
timeout 60 python tools/prof_conv.py 192 8 360 640
timeout 60 python tools/prof_conv.py 384 8 180 320
} > gpurun_out/s14_perf.txt 2>&1
cat gpurun_out/s14_perf.txt
timeout 200 python tools/vae_chunk_sweep.py 3 > gpurun_out/s14_vae.txt 2>&1
tail -2 gpurun_out/s14_vae.txt
