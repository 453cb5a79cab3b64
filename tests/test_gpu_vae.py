"""GPU parity of the Wan VAE decoder path: the implicit-GEMM convolution against F.conv3d with the reference's causal
padding, the phase-decomposed upsample convolution against Upsample(nearest-exact) + Conv2d, RMS_norm+SiLU against the
oracle, and the whole decoder against (a) the fixture produced by the REAL WanVAE_ on CPU in fp32 and (b) the oracle's
frame-by-frame restatement executed on the GPU.  The reference VAE computes in fp32 (TF32 on the GPU); this path keeps bf16
activations, so results are reported as PSNR with a stated floor, plus max-abs-error caps on the [-1, 1] output."""
import os

import pytest
import torch
import torch.nn.functional as F
from safetensors import safe_open

from oracle import vae_oracle as V
from oracle.wan_oracle import psnr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from lightx2v_b200 import lib as L

    L.load()
    return L


def _cl(x):      # [C, T, H, W] -> channels-last bf16 [T, H, W, C]
    return x.permute(1, 2, 3, 0).contiguous().to(torch.bfloat16)


@pytest.mark.parametrize("cin,cout,k,T,H,W", [(96, 96, (3, 3, 3), 3, 12, 40), (192, 96, (1, 1, 1), 2, 8, 32), (384, 384, (3, 3, 3), 2, 6, 33),
                                              (32, 384, (3, 3, 3), 3, 8, 8), (96, 16, (3, 3, 3), 2, 9, 70), (192, 384, (3, 1, 1), 4, 5, 32)])
def test_conv3d_cl_vs_causal_conv3d(lib, cin, cout, k, T, H, W):
    from lightx2v_b200.host.wan_vae import _Conv

    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(cin, T, H, W, generator=g, device="cuda")
    w = torch.randn(cout, cin, *k, generator=g, device="cuda") / (cin * k[0] * k[1] * k[2]) ** 0.5
    b = torch.randn(cout, generator=g, device="cuda") * 0.1
    xb, wb, bb = x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), b.to(torch.bfloat16).float()
    torch.backends.cudnn.allow_tf32 = False
    ref = V.causal_conv3d(xb.unsqueeze(0), wb, bb)[0]                       # fp32 math on the bf16-rounded operands
    conv = _Conv(w, b, "cuda")
    res = torch.randn(T, H, W, cout, generator=g, device="cuda").to(torch.bfloat16)
    out = conv(_cl(x))
    got = out.permute(3, 0, 1, 2).float()
    assert (got - ref).abs().max() <= 2e-2 + 1e-2 * ref.abs().max()
    assert psnr(got, ref) > 45
    out2 = conv(_cl(x), residual=res)
    ref2 = ref + res.permute(3, 0, 1, 2).float()
    assert psnr(out2.permute(3, 0, 1, 2).float(), ref2) > 45


def test_upsample_conv_phase_decomposition(lib):
    from lightx2v_b200.host.wan_vae import _UpsampleConv

    g = torch.Generator(device="cuda").manual_seed(2)
    C, T, H, W = 192, 2, 10, 36
    x = torch.randn(C, T, H, W, generator=g, device="cuda")
    w = torch.randn(C // 2, C, 3, 3, generator=g, device="cuda") / (C * 9) ** 0.5
    b = torch.randn(C // 2, generator=g, device="cuda") * 0.1
    xb = x.to(torch.bfloat16).float()
    torch.backends.cudnn.allow_tf32 = False
    frames = xb.permute(1, 0, 2, 3)
    ref = F.conv2d(F.interpolate(frames, scale_factor=(2.0, 2.0), mode="nearest-exact"), w, b, padding=1).permute(1, 0, 2, 3)
    got = _UpsampleConv(w, b, "cuda")(_cl(x)).permute(3, 0, 1, 2).float()
    assert got.shape == ref.shape
    assert psnr(got, ref) > 40          # pre-summed bf16 weights vs fp32 weights on the upsampled image


@pytest.mark.parametrize("C", [96, 192, 384])
def test_rms_silu(lib, C):
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(1000, C, generator=g, device="cuda").to(torch.bfloat16)
    gamma = 1 + 0.1 * torch.randn(C, generator=g, device="cuda")
    ref = F.silu(V.rms_norm(x.float(), gamma, channel_dim=1))
    got = lib.rms_silu_cl(x, gamma).float()
    assert (got - ref).abs().max() < 2e-2
    got2 = lib.rms_silu_cl(x, gamma, silu=False).float()
    assert (got2 - V.rms_norm(x.float(), gamma, channel_dim=1)).abs().max() < 3e-2


def test_decoder_vs_reference_fixture(golden_dir):
    from lightx2v_b200.host.wan_vae import WanVAEDecoderB200

    with safe_open(os.path.join(golden_dir, "wan_vae_decode_small.safetensors"), framework="pt") as f:
        zs, ref = f.get_tensor("zs"), f.get_tensor("images")
    dec = WanVAEDecoderB200(V.synth_vae_weights(0), device="cuda")
    out = dec.decode(zs.cuda()).cpu()
    assert out.shape == ref.shape
    p = psnr(out, ref)
    err = (out - ref).abs()
    print(f"VAE decode vs real reference (fp32 CPU): PSNR {p:.1f} dB, max abs err {err.max():.4f}, mean {err.mean():.5f}")
    assert p > 35 and err.max() < 0.15 and err.mean() < 1.5e-2


def test_decoder_vs_oracle_on_gpu_larger():
    """[16, 5, 24, 40] latent -> 17 frames of 192 x 320: whole-sequence bf16 path vs the frame-by-frame fp32/TF32 oracle on the GPU."""
    from lightx2v_b200.host.wan_vae import WanVAEDecoderB200

    W = V.synth_vae_weights(1, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(7)
    zs = torch.randn(16, 5, 24, 40, generator=g, device="cuda")
    ref = V.vae_decode(W, zs)
    out = WanVAEDecoderB200(W, device="cuda").decode(zs)
    assert out.shape == ref.shape == (1, 3, 17, 192, 320)
    p = psnr(out, ref)
    print(f"VAE decode 17x192x320 vs oracle on GPU: PSNR {p:.1f} dB, max abs err {(out - ref).abs().max():.4f}")
    assert p > 35


@pytest.mark.parametrize("shape,chunk", [((16, 7, 12, 16), 3), ((16, 5, 8, 33), 1), ((16, 4, 10, 12), 2)])
def test_chunked_decode_is_bit_identical_to_whole_sequence(shape, chunk):
    """Streaming decode (chunks of latent frames, every temporal convolution carrying the last two frames of its input: the reference's
    feat_cache protocol with a chunk size > 1) == the whole-sequence causal convolution, bit for bit; chunk = 1 is the reference's own
    one-latent-frame-per-iteration schedule (vae.py:723-737)."""
    from lightx2v_b200.host.wan_vae import WanVAEDecoderB200

    W = V.synth_vae_weights(2, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(3)
    zs = torch.randn(*shape, generator=g, device="cuda")
    whole = WanVAEDecoderB200(W, device="cuda", chunk_frames=None).decode(zs)
    chunked = WanVAEDecoderB200(W, device="cuda", chunk_frames=chunk).decode(zs)
    assert whole.shape == chunked.shape == (1, 3, 1 + 4 * (shape[1] - 1), shape[2] * 8, shape[3] * 8)
    assert torch.equal(whole, chunked), (whole - chunked).abs().max()


def test_bf16_decode_error_is_bounded_against_an_fp64_decode(record, golden_dir):
    """The reference decodes in fp32 (cuDNN: TF32 tensor cores by default); this path keeps bf16 activations.  Both are measured against a
    float64 evaluation of the same decoder ("truth") at the committed fixture's latent and at 17 x 192 x 320: PSNR and max error of
       (a) this bf16 path,  (b) the reference's GPU arithmetic (the frame-by-frame oracle in fp32 with TF32 convolutions),
    recorded to the parity-numbers file.  Floors: PSNR(bf16, truth) > 40 dB and max error < 0.1 on the [-1, 1] image."""
    from lightx2v_b200.host.wan_vae import WanVAEDecoderB200

    with safe_open(os.path.join(golden_dir, "wan_vae_decode_small.safetensors"), framework="pt") as f:
        zs_fix = f.get_tensor("zs").cuda()
    g = torch.Generator(device="cuda").manual_seed(7)
    cases = {"fixture": (0, zs_fix), "17x192x320": (1, torch.randn(16, 5, 24, 40, generator=g, device="cuda"))}
    for name, (seed, zs) in cases.items():
        W = V.synth_vae_weights(seed, device="cuda")
        W64 = {k: v.double() for k, v in W.items()}
        truth = V.vae_decode(W64, zs.double())
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        tf32 = V.vae_decode(W, zs.float())
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        fp32 = V.vae_decode(W, zs.float())
        ours = WanVAEDecoderB200(W, device="cuda").decode(zs)
        stats = {}
        for tag, t in (("bf16_b200", ours), ("reference_tf32", tf32), ("reference_fp32", fp32)):
            stats[tag + "_psnr_db"] = psnr(t, truth)
            stats[tag + "_max_err"] = (t.double() - truth.double()).abs().max().item()
        print(name, {k: round(v, 4) for k, v in stats.items()})
        record(**{f"{name}.{k}": v for k, v in stats.items()})
        assert stats["bf16_b200_psnr_db"] > 40 and stats["bf16_b200_max_err"] < 0.1, stats


def test_strided_and_temporal_downsample_convs_vs_torch(lib):
    """_DownConv (ZeroPad2d + Conv2d stride 2 through four parity views) and _TimeDownConv (3x1x1, stride 2 in time, first frame passed
    through) against torch on bf16-rounded operands."""
    from lightx2v_b200.host.wan_vae import _DownConv, _TimeDownConv

    g = torch.Generator(device="cuda").manual_seed(4)
    C, T, H, W = 96, 5, 14, 36
    x = torch.randn(C, T, H, W, generator=g, device="cuda")
    w = torch.randn(C, C, 3, 3, generator=g, device="cuda") / (C * 9) ** 0.5
    b = torch.randn(C, generator=g, device="cuda") * 0.1
    xb, wb, bb = x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), b.to(torch.bfloat16).float()
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(F.pad(xb.permute(1, 0, 2, 3), (0, 1, 0, 1)), wb, bb, stride=2).permute(1, 0, 2, 3)          # [C, T, H/2, W/2]
    got = _DownConv(w, b, "cuda")(_cl(x)).permute(3, 0, 1, 2).float()
    assert got.shape == ref.shape and psnr(got, ref) > 42, psnr(got, ref)
    wt = torch.randn(C, C, 3, 1, 1, generator=g, device="cuda") / (C * 3) ** 0.5
    wtb = wt.to(torch.bfloat16).float()
    y = F.conv3d(torch.cat([xb[:, :1], xb[:, 1:]], 1).unsqueeze(0), wtb, bb, stride=(2, 1, 1))[0]     # windows (0,1,2), (2,3,4)
    ref_t = torch.cat([xb[:, :1], y], dim=1)
    got_t = _TimeDownConv(wt, b, "cuda")(_cl(x)).permute(3, 0, 1, 2).float()
    assert got_t.shape == ref_t.shape == (C, 3, H, W)
    assert torch.equal(got_t[:, 0], xb[:, 0]) and psnr(got_t, ref_t) > 42, psnr(got_t, ref_t)


def test_encoder_vs_reference_fixture_and_oracle(golden_dir, record):
    """WanVAEEncoderB200 vs (a) the fixture of the REAL WanVAE_.encode (fp32 CPU) and (b) the oracle at [3, 9, 96, 160] on the GPU;
    bf16 activations vs the reference's fp32 -> PSNR of the latent mean, recorded."""
    from lightx2v_b200.host.wan_vae import WanVAEEncoderB200

    with safe_open(os.path.join(golden_dir, "wan_vae_encode_small.safetensors"), framework="pt") as f:
        video, mu = f.get_tensor("video"), f.get_tensor("mu")
    W = V.synth_vae_encoder_weights(0)
    out = WanVAEEncoderB200(W, device="cuda").encode(video.cuda()).cpu()
    assert out.shape == mu.shape
    p1 = psnr(out, mu)
    Wg = V.synth_vae_encoder_weights(1, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(8)
    vid = torch.rand(3, 9, 96, 160, generator=g, device="cuda") * 2 - 1
    ref = V.vae_encode(Wg, vid)
    got = WanVAEEncoderB200(Wg, device="cuda").encode(vid)
    p2 = psnr(got, ref)
    print(f"VAE encode: PSNR vs real-reference fixture {p1:.1f} dB, vs oracle at 9x96x160 {p2:.1f} dB")
    record(psnr_fixture_db=p1, psnr_oracle_db=p2, max_err_fixture=(out - mu).abs().max())
    assert p1 > 35 and p2 > 35


@pytest.mark.parametrize("cin,cout,T,H,W", [(96, 96, 3, 7, 640), (192, 192, 2, 6, 512), (96, 16, 2, 5, 700), (192, 96, 2, 4, 512), (384, 384, 2, 3, 512)])
def test_halo_staged_conv_vs_torch_and_vs_per_tap_tiles(lib, cin, cout, T, H, W):
    """csrc/conv3d_halo.cu (wide 3x3x3 stages: one staged halo tile, nine taps as row-shifted UMMA views) against fp32 torch on the
    bf16-rounded operands, with ragged H / W, the residual epilogue and the streaming (two-frame history) tap set - and against the per-tap
    tiles of conv3d.cu, which must agree to the last bit of accumulation-order noise (same products, same fp32 accumulator, different order
    of the 27 x cin terms)."""
    from lightx2v_b200.host.wan_vae import _Conv

    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(cin, T, H, W, generator=g, device="cuda")
    w = torch.randn(cout, cin, 3, 3, 3, generator=g, device="cuda") / (cin * 27) ** 0.5
    b = torch.randn(cout, generator=g, device="cuda") * 0.1
    xb, wb, bb = x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), b.to(torch.bfloat16).float()
    torch.backends.cudnn.allow_tf32 = False
    ref = V.causal_conv3d(xb.unsqueeze(0), wb, bb)[0]
    conv = _Conv(w, b, "cuda")
    res = torch.randn(T, H, W, conv.cout, generator=g, device="cuda").to(torch.bfloat16)
    buf = torch.zeros(T + 2, H, W, conv.cin, dtype=torch.bfloat16, device="cuda")
    buf[2:] = _cl(x)
    outs = {}
    prev = lib.get_option("conv_halo")
    try:
        for halo in (2, 0):                       # 2: halo tiles for every eligible width (the default, 1, uses them for the 192-wide tiles only)
            lib.set_option("conv_halo", halo)
            outs[1 if halo else 0] = (conv(_cl(x)).clone(), conv(_cl(x), residual=res).clone(), conv.causal(buf).clone())
    finally:
        lib.set_option("conv_halo", prev)
    got, got_r, got_h = (t.permute(3, 0, 1, 2).float()[:cout] for t in outs[1])
    assert psnr(got, ref) > 45 and (got - ref).abs().max() <= 2e-2 + 1e-2 * ref.abs().max()
    assert psnr(got_r, ref + res.permute(3, 0, 1, 2).float()[:cout]) > 45
    assert torch.equal(outs[1][0], outs[1][2])                         # history form == zero-padded form
    for a, b_ in zip(outs[1], outs[0]):                                # halo tiles vs per-tap tiles: bf16 outputs at most one ulp apart
        d = (a.float() - b_.float()).abs()
        assert d.max() <= 2.0 ** -6 * max(1.0, b_.float().abs().max().item()) and (d > 0).float().mean() < 0.05
