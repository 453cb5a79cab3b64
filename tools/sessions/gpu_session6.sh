#!/bin/bash
# 2-GPU session: tile-parallel Hunyuan VAE after the plan/compute/exchange/blend restructure + Hunyuan N=2 bench line
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest 2-GPU"; timeout 900 python -m pytest tests/test_gpu_ulysses.py -m gpu -v -rs > gpurun_out/s6_pytest_2gpu.txt 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/s6_pytest_2gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "== bench N=2 hunyuan"; timeout 600 $TR bench.py --gpus 2 --steps 2 --warmup 3 --workload hunyuan-13b-720p-129f > gpurun_out/s6_bench_n2_hunyuan.jsonl 2> gpurun_out/s6_bench_n2_hunyuan.err; echo "rc=$?"; tail -c 1600 gpurun_out/s6_bench_n2_hunyuan.jsonl; grep "bench +" gpurun_out/s6_bench_n2_hunyuan.err | tail -5
