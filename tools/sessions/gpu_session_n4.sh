#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519"
echo "== bench N=4 t2v"; timeout 400 $TR bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/s7_bench_n4_t2v.jsonl 2> gpurun_out/s7_bench_n4_t2v.err; echo "rc=$?"; tail -c 900 gpurun_out/s7_bench_n4_t2v.jsonl; grep "bench +" gpurun_out/s7_bench_n4_t2v.err | tail -4
echo "== bench N=4 hunyuan"; timeout 400 $TR bench.py --gpus 4 --steps 3 --warmup 3 --workload hunyuan-13b-720p-129f > gpurun_out/s7_bench_n4_hunyuan.jsonl 2> gpurun_out/s7_bench_n4_hunyuan.err; echo "rc=$?"; tail -c 900 gpurun_out/s7_bench_n4_hunyuan.jsonl
