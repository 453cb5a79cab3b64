"""Parity + timing of the halo-staged convolution (csrc/conv3d_halo.cu, B200_CONV_HALO=1) against torch on bf16-rounded operands, at the
shapes it serves: 96->96 and 192->192 (and 96->16, 192->96... where eligible), W >= 512, with ragged H / W, residual and the hist taps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from lightx2v_b200 import lib  # noqa: E402
from lightx2v_b200.host.wan_vae import _Conv  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
g = torch.Generator(device="cuda").manual_seed(0)


def psnr(a, b):
    return float(10 * torch.log10(b.abs().max() ** 2 / (a - b).pow(2).mean().clamp(min=1e-30)))


def cl(x):
    return x.permute(1, 2, 3, 0).contiguous().to(torch.bfloat16)


for cin, cout, T, H, W in ((96, 96, 3, 7, 640), (192, 192, 2, 6, 512), (96, 16, 2, 5, 700), (192, 96, 2, 4, 512)):
    x = torch.randn(cin, T, H, W, generator=g, device="cuda")
    w = torch.randn(cout, cin, 3, 3, 3, generator=g, device="cuda") / (cin * 27) ** 0.5
    b = torch.randn(cout, generator=g, device="cuda") * 0.1
    xb, wb, bb = x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), b.to(torch.bfloat16).float()
    ref = F.conv3d(F.pad(xb.unsqueeze(0), (1, 1, 1, 1, 2, 0)), wb, bb)[0]
    conv = _Conv(w, b, "cuda")
    res = torch.randn(T, H, W, conv.cout, generator=g, device="cuda").to(torch.bfloat16)
    got = conv(cl(x)).permute(3, 0, 1, 2).float()[:cout]
    got_r = conv(cl(x), residual=res).permute(3, 0, 1, 2).float()[:cout]
    ref_r = ref + res.permute(3, 0, 1, 2).float()[:cout]
    # streaming form: two zero history frames in front, taps shifted by +2
    buf = torch.zeros(T + 2, H, W, conv.cin, dtype=torch.bfloat16, device="cuda")
    buf[2:] = cl(x)
    got_h = conv.causal(buf).permute(3, 0, 1, 2).float()[:cout]
    print(f"{cin}->{cout} [{T},{H},{W}]: PSNR {psnr(got, ref):.1f} dB  max err {(got - ref).abs().max():.4f} | +residual {psnr(got_r, ref_r):.1f} dB | hist taps {psnr(got_h, ref):.1f} dB",
          flush=True)
