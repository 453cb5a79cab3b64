import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightx2v_b200 import lib
S, H = 75600, 40
qkv = torch.randn(S, 3, H, 128, device="cuda").bfloat16()
o = torch.empty(S, H, 128, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    lib.fmha(qkv[:, 0], qkv[:, 1], qkv[:, 2], out=o)
torch.cuda.synchronize()
