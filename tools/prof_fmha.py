import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightx2v_b200 import lib
S, H = 75600, 5
q = torch.randn(S, H, 128, device="cuda").bfloat16(); k = torch.randn(S, H, 128, device="cuda").bfloat16(); v = torch.randn(S, H, 128, device="cuda").bfloat16()
o = torch.empty_like(q)
for _ in range(3):
    lib.fmha(q, k, v, out=o)
torch.cuda.synchronize()
