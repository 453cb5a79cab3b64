#!/bin/bash
# ncu --set full of the two-issuer halo kernels at the decoder-stage shapes
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3d_halo -s 2 -c 1 -o gpurun_out/r02_halo96_v2 -f python tools/prof_conv.py 96 8 720 1280 > gpurun_out/s13_ncu_halo96.log 2>&1; echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3d_halo -s 2 -c 1 -o gpurun_out/r02_halo192_v2 -f python tools/prof_conv.py 192 8 360 640 > gpurun_out/s13_ncu_halo192.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/s13_ncu_halo96.log gpurun_out/s13_ncu_halo192.log
