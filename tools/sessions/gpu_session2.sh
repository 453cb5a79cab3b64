#!/bin/bash
# 1-GPU session: VAE / CogVideoX / drop-in tests, streaming-kernel rates, VAE decode timing + launch list (new conv tiles vs B200_CONV_NARROW=1)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest (vae, hunyuan vae, cogvideox, dropin, kernels)"; timeout 900 python -m pytest tests/test_gpu_vae.py tests/test_gpu_hunyuan_vae.py tests/test_gpu_cogvideox.py tests/test_gpu_reference_dropin.py tests/test_gpu_kernels.py -m gpu -q -rs > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/s2_pytest.log
echo "== perf_stream"; timeout 300 python tools/perf_stream.py > gpurun_out/s2_perf_stream.log 2>&1; echo "rc=$?"; cat gpurun_out/s2_perf_stream.log | cut -c1-200
echo "== VAE decode (new tiles)"; timeout 300 python tools/perf_vae.py > gpurun_out/s2_vae_new.json 2> gpurun_out/s2_vae_new.err; echo "rc=$?"; cat gpurun_out/s2_vae_new.json
echo "== VAE decode (round-1 narrow tiles)"; B200_CONV_NARROW=1 timeout 300 python tools/perf_vae.py > gpurun_out/s2_vae_narrow.json 2> gpurun_out/s2_vae_narrow.err; echo "rc=$?"; cat gpurun_out/s2_vae_narrow.json
echo "== VAE launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "vae/" --csv --log-file gpurun_out/s2_vae_launches.csv python tools/vae_decode_once.py 21 > gpurun_out/s2_vae_ncu.log 2>&1; echo "rc=$?"
python tools/launch_share.py gpurun_out/s2_vae_launches.csv gpurun_out/s2_vae_launch_shares.txt | head -30
