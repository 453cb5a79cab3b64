"""A few launches of one VAE convolution at a decoder-stage shape, for `ncu --set full -k regex:conv3d_igemm -s 2 -c 1`, and a CUDA-event timing.
usage: prof_conv.py C T H W [cout]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lightx2v_b200.host.wan_vae import _Conv  # noqa: E402

C, T, H, W = (int(v) for v in sys.argv[1:5])
cout = int(sys.argv[5]) if len(sys.argv) > 5 else C
g = torch.Generator(device="cuda").manual_seed(0)
pitch = int(os.environ.get("PITCH", "0"))        # channel pitch of the activation tensors in memory (0: dense); e.g. 128 for the 96-channel stage
if pitch:
    x = torch.randn(T, H, W, pitch, generator=g, device="cuda").to(torch.bfloat16)[..., :C]
else:
    x = torch.randn(T, H, W, C, generator=g, device="cuda").to(torch.bfloat16)
w = torch.randn(cout, C, 3, 3, 3, generator=g, device="cuda") / (C * 27) ** 0.5
conv = _Conv(w, torch.zeros(cout), "cuda")
out = torch.empty(T, H, W, pitch if pitch else conv.cout, dtype=torch.bfloat16, device="cuda")[..., :conv.cout]
for _ in range(3):
    conv(x, out=out)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
n = 5
for _ in range(n):
    conv(x, out=out)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / n
fl = 2.0 * T * H * W * 27 * C * cout
print(f"pitch {pitch} conv {C}->{cout} [{T},{H},{W}] mode env B200_CONV_NARROW={os.environ.get('B200_CONV_NARROW', '0')}: {ms:.3f} ms, {fl / ms / 1e9:.1f} TFLOP/s")
