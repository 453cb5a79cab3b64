#!/usr/bin/env python
"""One Wan VAE decode of the bench latent inside the NVTX range "vae" (for `ncu --nvtx --nvtx-include "vae/"` launch lists)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lightx2v_b200.host.wan_vae import WanVAEDecoderB200  # noqa: E402
from oracle import vae_oracle as V   # synthetic weights only  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 21
dec = WanVAEDecoderB200(V.synth_vae_weights(0), device="cuda")
zs = torch.randn(16, frames, 90, 160, device="cuda")
dec.decode(zs[:, :2, :16, :16])
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("vae")
dec.decode(zs)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
