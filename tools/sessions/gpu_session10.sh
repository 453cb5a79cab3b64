#!/bin/bash
# two MMA-issuing warps: parity of every conv path, then timing at the decoder-stage shapes and the whole decode
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_vae.py -x -q -m gpu > gpurun_out/s10_vae_tests.txt 2>&1
tail -5 gpurun_out/s10_vae_tests.txt
{
timeout 60 python tools/prof_conv.py 96 8 720 1280
timeout 60 python tools/prof_conv.py 96 8 720 1280 16
timeout 60 python tools/prof_conv.py 192 8 360 640
B200_CONV_HALO=0 timeout 60 python tools/prof_conv.py 192 8 360 640
timeout 60 python tools/prof_conv.py 384 8 180 320
timeout 60 python tools/prof_conv.py 192 8 720 1280 96
} > gpurun_out/s10_conv.txt 2>&1
cat gpurun_out/s10_conv.txt
timeout 200 python tools/vae_decode_once.py > gpurun_out/s10_vae_decode.txt 2>&1
tail -5 gpurun_out/s10_vae_decode.txt
