#!/bin/bash
# round-2 closing session on one GPU: whole GPU suite, the default bench line, the config-3 lines (fp8 / nvfp4 distilled sampler),
# head-conv A/B (halo tiles vs per-tap tiles at N = 16), VAE decode
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -q -m gpu > gpurun_out/s15_pytest.txt 2>&1; echo "rc=$?"; tail -4 gpurun_out/s15_pytest.txt
echo "== bench default"; timeout 420 python bench.py --budget-s 330 > gpurun_out/s15_bench_n1.jsonl 2> gpurun_out/s15_bench_n1.err; echo "rc=$?"; tail -c 2500 gpurun_out/s15_bench_n1.jsonl
echo "== head conv A/B"
{ B200_CONV_HALO=2 timeout 60 python tools/prof_conv.py 96 8 720 1280 16; timeout 60 python tools/prof_conv.py 96 8 720 1280 16; } > gpurun_out/s15_conv16.txt 2>&1; cat gpurun_out/s15_conv16.txt
echo "== VAE"; timeout 120 python tools/vae_chunk_sweep.py 3 > gpurun_out/s15_vae.txt 2>&1; tail -1 gpurun_out/s15_vae.txt
B200_CONV_HALO=2 timeout 120 python tools/vae_chunk_sweep.py 3 > gpurun_out/s15_vae_halo2.txt 2>&1; tail -1 gpurun_out/s15_vae_halo2.txt
echo "== bench fp8 distill"; timeout 240 python bench.py --workload wan2.1-t2v-14b-fp8-distill-720p-81f --no-cpu-baseline --no-gpu-reference --no-vae --budget-s 200 > gpurun_out/s15_bench_fp8.jsonl 2> gpurun_out/s15_bench_fp8.err; echo "rc=$?"; tail -c 700 gpurun_out/s15_bench_fp8.jsonl
echo "== bench nvfp4 distill"; timeout 240 python bench.py --workload wan2.1-t2v-14b-nvfp4-distill-720p-81f --no-cpu-baseline --no-gpu-reference --no-vae --budget-s 200 > gpurun_out/s15_bench_nvfp4.jsonl 2> gpurun_out/s15_bench_nvfp4.err; echo "rc=$?"; tail -c 700 gpurun_out/s15_bench_nvfp4.jsonl
