"""CUDA-event timing of the attention kernel at a bench shape.  usage: time_fmha.py S H D [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lightx2v_b200 import lib  # noqa: E402

S, H, D = (int(v) for v in sys.argv[1:4])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
qkv = torch.randn(S, 3, H, D, device="cuda").bfloat16()
o = torch.empty(S, H, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    lib.fmha(qkv[:, 0], qkv[:, 1], qkv[:, 2], out=o)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(iters):
    lib.fmha(qkv[:, 0], qkv[:, 1], qkv[:, 2], out=o)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / iters
print(f"fmha S={S} H={H} D={D}: {ms:.3f} ms, {4.0 * S * S * H * D / ms / 1e9:.1f} TFLOP/s", flush=True)
