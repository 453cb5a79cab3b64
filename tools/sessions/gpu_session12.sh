#!/bin/bash
mkdir -p gpurun_out
{
B200_CONV_HALO=2 timeout 60 python tools/prof_conv.py 96 8 720 1280
B200_CONV_HALO=2 timeout 60 python tools/prof_conv.py 96 8 720 1280 16
B200_CONV_HALO=2 timeout 60 python tools/prof_conv.py 192 8 720 1280 96
B200_CONV_HALO=2 timeout 60 python tools/prof_conv.py 192 8 360 640
} > gpurun_out/s12_conv_halo2.txt 2>&1
cat gpurun_out/s12_conv_halo2.txt
B200_CONV_HALO=2 timeout 200 python tools/vae_chunk_sweep.py 3 > gpurun_out/s12_vae_halo2.txt 2>&1
tail -4 gpurun_out/s12_vae_halo2.txt
timeout 200 python tools/vae_chunk_sweep.py 3 > gpurun_out/s12_vae_default.txt 2>&1
tail -4 gpurun_out/s12_vae_default.txt
