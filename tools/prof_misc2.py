#!/usr/bin/env python
"""Second ncu pass: the wide convolution tiles of the HunyuanVideo decoder (BLOCK_N = 128 / 256) and the 128x256 NVFP4 GEMM tile.
    ncu --set full --clock-control none --import-source on -k regex:'conv3d|gemm_bf16' -c 3 -o gpurun_out/r01_misc2 python tools/prof_misc2.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lightx2v_b200 import lib
from lightx2v_b200.host.hunyuan_vae import PAD, _conv333

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device=dev) * sc).to(torch.bfloat16)   # noqa: E731
x128 = lib.gn_apply_pad_cl(rnd(65, 256, 256, 128), None, None, None, pad=PAD)
c128 = _conv333(rnd(128, 128, 3, 3, 3, sc=0.02).float(), rnd(128).float(), dev)
x256 = lib.gn_apply_pad_cl(rnd(33, 128, 128, 256), None, None, None, pad=PAD)
c256 = _conv333(rnd(256, 256, 3, 3, 3, sc=0.02).float(), rnd(256).float(), dev)
S, D = 75600, 5120
a, w, b = rnd(S, D), rnd(D, D, sc=0.02), rnd(D)
gw, _ = lib.nvfp4_act_scale(w)
w4, sw4 = lib.quant_nvfp4(w, gw)
ga, alpha = lib.nvfp4_act_scale(a, gw)
a4, sa4 = lib.quant_nvfp4(a, ga)
o = torch.empty(S, D, device=dev, dtype=torch.bfloat16)
torch.cuda.synchronize()
c128(x128)                                                                  # 1 conv3d BLOCK_N = 128 (65 x 256 x 256 voxels)
c256(x256)                                                                  # 2 conv3d BLOCK_N = 256 (33 x 128 x 128 voxels)
lib.gemm_nvfp4(a4, w4, sa4, sw4, alpha, b, out=o, block_n=256)              # 3 nvfp4 GEMM, 128 x 256 tile
torch.cuda.synchronize()
print("done")
