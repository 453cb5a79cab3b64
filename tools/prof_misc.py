#!/usr/bin/env python
"""One launch of every non-FMHA kernel at its bench shape (14B / 720p x 81f; VAE stage shapes), for a single `ncu --set full` pass:
    ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16|ln_modulate|rms_rope|quant_|absmax|conv3d|gn_|rms_silu' \\
        -c 16 -o gpurun_out/r01_misc python tools/prof_misc.py
Each kernel is launched exactly once after the allocations, so the capture order below is the order in the report."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lightx2v_b200 import lib
from lightx2v_b200.host.hunyuan_vae import PAD, _conv333
from lightx2v_b200.host.wan_vae import _Conv

dev = "cuda"
S, D, F_ = 75600, 5120, 13824
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device=dev) * sc).to(torch.bfloat16)   # noqa: E731

x = rnd(S, D)
w_qkv, b_qkv = rnd(3 * D, D, sc=0.02), rnd(3 * D)
w_o, b_o, gate = rnd(D, D, sc=0.02), rnd(D), rnd(D)
qkv = torch.empty(S, 3 * D, device=dev, dtype=torch.bfloat16)
scale, shift = rnd(D, sc=0.1), rnd(D, sc=0.1)
n = torch.empty_like(x)
wq, wk = rnd(D), rnd(D)
cs = torch.rand(S, 64, 2, device=dev)
torch.cuda.synchronize()

# ---- DiT row-wise + GEMM kernels (bf16 / fp8 / nvfp4)
lib.ln_modulate(x, scale=scale, shift=shift, out=n)                                     # 1 ln_modulate
lib.gemm_bf16(n, w_qkv, b_qkv, out=qkv)                                                 # 2 gemm bf16 (QKV, N = 15360)
lib.rms_rope_(qkv[:, :D], wq, qkv[:, D:2 * D], wk, cos_sin=cs, rope_rows=S)              # 3 rms_rope
lib.gemm_bf16(n, w_o, b_o, out=x, epilogue=lib.EPI_GATE_RESIDUAL, gate=gate)            # 4 gemm bf16 (o-proj, gate residual epilogue)
aq, sa = lib.quant_fp8_per_token(n)                                                     # 5 quant fp8
wq8, sw8 = lib.quant_fp8_per_token(w_o)                                                 # 6 (weights)
lib.gemm_fp8(aq, sa, wq8, sw8, b_o, out=n)                                              # 7 gemm fp8
gw, _ = lib.nvfp4_act_scale(w_o)                                                        # 8 absmax
w4, sw4 = lib.quant_nvfp4(w_o, gw)                                                      # 9 quant nvfp4 (weights)
ga, alpha = lib.nvfp4_act_scale(x, gw)                                                  # 10 absmax (activations)
a4, sa4 = lib.quant_nvfp4(x, ga)                                                        # 11 quant nvfp4 (activations)
lib.gemm_nvfp4(a4, w4, sa4, sw4, alpha, b_o, out=n)                                     # 12 gemm nvfp4
torch.cuda.synchronize()
del qkv, w_qkv

# ---- VAE kernels: Wan stage 3 (96 ch @ 8 frames of 720 x 1280) and Hunyuan stage 3 (128 ch, one tile 65 x 256 x 256)
xw = rnd(8, 720, 1280, 96)
cw = _Conv(rnd(96, 96, 3, 3, 3, sc=0.02).float(), rnd(96).float(), dev)
gam = torch.ones(96, device=dev)
yw = lib.rms_silu_cl(xw, gam)                                                           # 13 rms_silu
cw(yw)                                                                                  # 14 conv3d BLOCK_N = 96
del xw, yw
xh = rnd(65, 256, 256, 128)
ch = _conv333(rnd(128, 128, 3, 3, 3, sc=0.02).float(), rnd(128).float(), dev)
sums = lib.gn_stats_cl(xh)                                                              # 15 gn_stats
xp = lib.gn_apply_pad_cl(xh, sums, torch.ones(128, device=dev), torch.zeros(128, device=dev), pad=PAD)   # 16 gn_apply_pad
ch(xp)                                                                                  # 17 conv3d BLOCK_N = 128 (wide)
torch.cuda.synchronize()
print("done")
