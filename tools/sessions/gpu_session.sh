#!/bin/bash
# One gpurun call = one packed session: tests, bench, launch lists.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L > gpurun_out/s_gpus.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q -rs > gpurun_out/s_pytest_gpu.log 2>&1; echo "pytest rc=$?" ; tail -5 gpurun_out/s_pytest_gpu.log
echo "== perf_stream" ; timeout 300 python tools/perf_stream.py > gpurun_out/s_perf_stream.log 2>&1; echo "rc=$?"; tail -20 gpurun_out/s_perf_stream.log
echo "== bench N=1 (default schedule)" ; timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/s_bench_n1.jsonl 2> gpurun_out/s_bench_n1.err; echo "rc=$?"; tail -c 1500 gpurun_out/s_bench_n1.jsonl; tail -5 gpurun_out/s_bench_n1.err
echo "== VAE launch list" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "vae/" --csv --log-file gpurun_out/s_vae_launches.csv python tools/vae_decode_once.py 21 > gpurun_out/s_vae_ncu.log 2>&1; echo "rc=$?"
python tools/launch_share.py gpurun_out/s_vae_launches.csv gpurun_out/s_vae_launch_shares.txt | head -30
