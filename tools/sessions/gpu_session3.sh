#!/bin/bash
# 2-GPU session: every 2-GPU test + N = 2 bench lines (t2v fused Ulysses, i2v, HunyuanVideo blocks + VAE)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L
echo "== pytest 2-GPU"; timeout 900 python -m pytest tests/test_gpu_ulysses.py -m gpu -v -rs > gpurun_out/s3_pytest_2gpu.txt 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/s3_pytest_2gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "== bench N=2 t2v"; timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/s3_bench_n2_t2v.jsonl 2> gpurun_out/s3_bench_n2_t2v.err; echo "rc=$?"; tail -c 1200 gpurun_out/s3_bench_n2_t2v.jsonl; grep "bench +" gpurun_out/s3_bench_n2_t2v.err | tail -8
echo "== bench N=2 i2v"; timeout 600 $TR bench.py --gpus 2 --steps 2 --warmup 3 --workload wan2.1-i2v-14b-720p-81f > gpurun_out/s3_bench_n2_i2v.jsonl 2> gpurun_out/s3_bench_n2_i2v.err; echo "rc=$?"; tail -c 600 gpurun_out/s3_bench_n2_i2v.jsonl
echo "== bench N=2 hunyuan"; timeout 600 $TR bench.py --gpus 2 --steps 2 --warmup 3 --workload hunyuan-13b-720p-129f > gpurun_out/s3_bench_n2_hunyuan.jsonl 2> gpurun_out/s3_bench_n2_hunyuan.err; echo "rc=$?"; tail -c 1500 gpurun_out/s3_bench_n2_hunyuan.jsonl; tail -5 gpurun_out/s3_bench_n2_hunyuan.err
