#!/bin/bash
# 1-GPU session: full GPU suite, conv tile A/B + ncu, VAE chunk sweep, graph bench, FMHA ncu capture
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu (full)"; timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/s4_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s4_pytest.log
echo "== UMMA row-shift probe"; timeout 120 python tools/probe_rowshift.py 2>&1 | tee gpurun_out/s4_probe_rowshift.txt | tail -6
echo "== halo conv parity + timing"
for bo in 1 0; do
  echo "-- B200_CONV_HALO=1 B200_HALO_BASE_OFFSET=$bo"
  B200_CONV_HALO=1 B200_HALO_BASE_OFFSET=$bo timeout 300 python tools/check_halo_conv.py 2>&1 | tail -8
done | tee gpurun_out/s4_halo.txt
echo "== conv A/B"
for args in "96 8 720 1280" "192 8 360 640" "384 8 180 320" "192 8 360 640 96" "96 8 720 1280 16"; do
  timeout 120 python tools/prof_conv.py $args 2>&1 | tail -1
  B200_CONV_NARROW=1 timeout 120 python tools/prof_conv.py $args 2>&1 | tail -1
done | tee gpurun_out/s4_conv_ab.txt
echo "== VAE chunk sweep"; timeout 300 python tools/vae_chunk_sweep.py 2 3 5 7 2>&1 | tee gpurun_out/s4_vae_chunks.jsonl | tail -5
echo "== ncu conv<96,2>"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3d_igemm -s 2 -c 1 -o gpurun_out/r02_conv96 python tools/prof_conv.py 96 8 720 1280 > gpurun_out/s4_ncu_conv96.log 2>&1; echo "rc=$?"
echo "== ncu conv<192,1>"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3d_igemm -s 2 -c 1 -o gpurun_out/r02_conv192 python tools/prof_conv.py 192 8 360 640 > gpurun_out/s4_ncu_conv192.log 2>&1; echo "rc=$?"
echo "== ncu fmha 75600x40"; timeout 600 ncu --set full --clock-control none -k regex:fmha_fwd_kernel -s 2 -c 1 -o gpurun_out/r02_fmha_h40 python tools/prof_fmha_shapes.py 75600 40 128 > gpurun_out/s4_ncu_fmha.log 2>&1; echo "rc=$?"
echo "== bench --graph"; timeout 600 python bench.py --steps 3 --warmup 3 --graph --no-vae --no-gpu-reference --no-cpu-baseline > gpurun_out/s4_bench_graph.jsonl 2> gpurun_out/s4_bench_graph.err; echo "rc=$?"; tail -c 700 gpurun_out/s4_bench_graph.jsonl; grep "bench +" gpurun_out/s4_bench_graph.err | tail -4
ls -la gpurun_out/*.ncu-rep | tail -5
