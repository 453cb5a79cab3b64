#!/bin/bash
# N-GPU session (N = $1, default 8): bench lines for BASELINE configs 2 (t2v), 4 (i2v) and 5 (HunyuanVideo blocks + VAE) at N ranks
set -u
N=${1:-8}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
echo "== bench N=$N t2v"; timeout 500 $TR bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/s8_bench_n${N}_t2v.jsonl 2> gpurun_out/s8_bench_n${N}_t2v.err; echo "rc=$?"; tail -c 1500 gpurun_out/s8_bench_n${N}_t2v.jsonl; grep "bench +" gpurun_out/s8_bench_n${N}_t2v.err | tail -6
echo "== bench N=$N i2v"; timeout 500 $TR bench.py --gpus $N --steps 4 --warmup 3 --workload wan2.1-i2v-14b-720p-81f > gpurun_out/s8_bench_n${N}_i2v.jsonl 2> gpurun_out/s8_bench_n${N}_i2v.err; echo "rc=$?"; tail -c 800 gpurun_out/s8_bench_n${N}_i2v.jsonl
echo "== bench N=$N hunyuan"; timeout 500 $TR bench.py --gpus $N --steps 4 --warmup 3 --workload hunyuan-13b-720p-129f > gpurun_out/s8_bench_n${N}_hunyuan.jsonl 2> gpurun_out/s8_bench_n${N}_hunyuan.err; echo "rc=$?"; tail -c 1500 gpurun_out/s8_bench_n${N}_hunyuan.jsonl; tail -3 gpurun_out/s8_bench_n${N}_hunyuan.err
echo "== bench N=$N t2v cfg-parallel"; timeout 500 $TR bench.py --gpus $N --steps 4 --warmup 3 --parallel cfg > gpurun_out/s8_bench_n${N}_t2v_cfg.jsonl 2> gpurun_out/s8_bench_n${N}_t2v_cfg.err; echo "rc=$?"; tail -c 600 gpurun_out/s8_bench_n${N}_t2v_cfg.jsonl
