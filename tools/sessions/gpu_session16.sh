#!/bin/bash
# re-validation after the last two changes (halo tiles for the 16-wide head conv by default; Hunyuan i2v test assertion) + final ncu captures
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 400 python -m pytest tests/test_gpu_hunyuan.py tests/test_gpu_vae.py tests/test_gpu_hunyuan_vae.py -q -m gpu > gpurun_out/s16_pytest.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/s16_pytest.txt
echo "== VAE"; timeout 120 python tools/vae_chunk_sweep.py 3 > gpurun_out/s16_vae.txt 2>&1; tail -1 gpurun_out/s16_vae.txt
echo "== ncu fmha"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:fmha_fwd_kernel -s 2 -c 1 -f -o gpurun_out/r02_fmha_h40_v2 python tools/prof_fmha_shapes.py 75600 40 128 > gpurun_out/s16_ncu_fmha.log 2>&1; echo "rc=$?"
echo "== ncu halo96"; timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv3d_halo -s 2 -c 1 -f -o gpurun_out/r02_halo96_v3 python tools/prof_conv.py 96 8 720 1280 > gpurun_out/s16_ncu_halo96.log 2>&1; echo "rc=$?"
echo "== VAE launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "vae/" --csv --log-file gpurun_out/s16_vae_launches.csv python tools/vae_decode_once.py 21 > gpurun_out/s16_vae_ncu.log 2>&1; echo "rc=$?"
