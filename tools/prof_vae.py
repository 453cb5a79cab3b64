import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightx2v_b200.host.wan_vae import WanVAEDecoderB200
from oracle import vae_oracle as V   # synthetic weights only
dec = WanVAEDecoderB200(V.synth_vae_weights(0), device="cuda")
zs = torch.randn(16, 4, 90, 160, device="cuda")
for _ in range(2):
    dec.decode(zs)
torch.cuda.synchronize()
