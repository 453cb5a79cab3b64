"""Two warm launches + one profiled launch of the attention kernel at a bench shape: `ncu --set full -k regex:fmha_fwd_kernel -s 2 -c 1`.
usage: prof_fmha_shapes.py S H D"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lightx2v_b200 import lib  # noqa: E402

S, H, D = (int(v) for v in sys.argv[1:4])
qkv = torch.randn(S, 3, H, D, device="cuda").bfloat16()
o = torch.empty(S, H, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    lib.fmha(qkv[:, 0], qkv[:, 1], qkv[:, 2], out=o)
torch.cuda.synchronize()
