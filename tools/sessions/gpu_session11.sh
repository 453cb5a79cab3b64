#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tools/probe_umma_rate.py > gpurun_out/s11_umma_rate.txt 2>&1
cat gpurun_out/s11_umma_rate.txt
