#!/bin/bash
# 1-GPU session: halo conv default ON - VAE tests, conv timings halo vs per-tap, VAE decode bench, launch list, ncu of the halo kernel
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest VAE"; timeout 900 python -m pytest tests/test_gpu_vae.py tests/test_gpu_hunyuan_vae.py -m gpu -q -rs > gpurun_out/s5_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/s5_pytest.log
echo "== conv timing halo on/off"
for args in "96 8 720 1280" "192 8 360 640" "96 8 720 1280 16" "192 8 360 640 96"; do
  B200_CONV_HALO=1 timeout 120 python tools/prof_conv.py $args 2>&1 | tail -1 | sed 's/$/ [halo]/'
  B200_CONV_HALO=0 timeout 120 python tools/prof_conv.py $args 2>&1 | tail -1 | sed 's/$/ [per-tap]/'
done | tee gpurun_out/s5_conv_halo_ab.txt
echo "== VAE decode"; timeout 300 python tools/perf_vae.py > gpurun_out/s5_vae_halo.json 2> gpurun_out/s5_vae_halo.err; echo "rc=$?"; cat gpurun_out/s5_vae_halo.json
echo "== VAE launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "vae/" --csv --log-file gpurun_out/s5_vae_launches.csv python tools/vae_decode_once.py 21 > gpurun_out/s5_vae_ncu.log 2>&1; echo "rc=$?"
python tools/launch_share.py gpurun_out/s5_vae_launches.csv gpurun_out/s5_vae_launch_shares.txt | head -24
echo "== ncu halo<96>"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3d_halo -s 2 -c 1 -o gpurun_out/r02_halo96 python tools/prof_conv.py 96 8 720 1280 > gpurun_out/s5_ncu_halo96.log 2>&1; echo "rc=$?"
echo "== ncu halo<192>"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3d_halo -s 2 -c 1 -o gpurun_out/r02_halo192 python tools/prof_conv.py 192 8 360 640 > gpurun_out/s5_ncu_halo192.log 2>&1; echo "rc=$?"
