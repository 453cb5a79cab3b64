"""ORACLE (test infrastructure, NOT product code): functional restatement of the Wan 3D causal VAE *decoder* exactly as the
reference executes it — one latent frame per iteration with a two-frame causal cache per convolution
(lightx2v/models/video_encoders/hf/wan/vae.py: CausalConv3d :19-44, RMS_norm :47-59, Resample :70-159, ResidualBlock :185-223,
AttentionBlock :226-262, Decoder3d :377-489, WanVAE_.decode :713-738, WanVAE.decode :931-957).

Weights are a flat dict with the reference's state_dict key names (`conv2.*`, `decoder.conv1.*`, `decoder.middle.{0,1,2}.*`,
`decoder.upsamples.N.*`, `decoder.head.{0,2}.*`).  Pinned by fixtures generated from the REAL `WanVAE_` class
(oracle/gen_golden.py:gen_vae_fixture -> tests/golden/wan_vae_decode_small.safetensors; tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

CACHE_T = 2  # vae.py:16

MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]
STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]   # vae.py:804-839


def decoder_layout(dim=96, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temperal_upsample=(True, True, False)):
    """Module list of Decoder3d.upsamples (vae.py:409-427): [("res", in, out) | ("up2d"/"up3d", dim)], plus dims[0]."""
    dims = [dim * u for u in [dim_mult[-1]] + list(dim_mult[::-1])]
    layers = []
    for i, (in_dim, out_dim) in enumerate(zip(dims[:-1], dims[1:])):
        if i in (1, 2, 3):
            in_dim = in_dim // 2
        for _ in range(num_res_blocks + 1):
            layers.append(("res", in_dim, out_dim))
            in_dim = out_dim
        if i != len(dim_mult) - 1:
            layers.append(("up3d" if temperal_upsample[i] else "up2d", out_dim))
    return dims[0], layers, out_dim


def synth_vae_weights(seed=0, dim=96, z_dim=16, device="cpu", dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded random decoder weights under the reference's state_dict names (realistic fan-in scaling so activations stay O(1))."""
    g = torch.Generator().manual_seed(seed)
    W: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k):
        fan = cin * math.prod(k)
        W[name + ".weight"] = (torch.randn(cout, cin, *k, generator=g) * (1.0 / math.sqrt(fan))).to(dtype).to(device)
        W[name + ".bias"] = (torch.randn(cout, generator=g) * 0.05).to(dtype).to(device)

    def gamma(name, c, images):
        shape = (c, 1, 1) if images else (c, 1, 1, 1)
        W[name] = (1.0 + 0.1 * torch.randn(shape, generator=g)).to(dtype).to(device)

    def res(prefix, cin, cout):
        gamma(prefix + ".residual.0.gamma", cin, False)
        conv(prefix + ".residual.2", cout, cin, (3, 3, 3))
        gamma(prefix + ".residual.3.gamma", cout, False)
        conv(prefix + ".residual.6", cout, cout, (3, 3, 3))
        if cin != cout:
            conv(prefix + ".shortcut", cout, cin, (1, 1, 1))

    d0, layers, d_out = decoder_layout(dim, z_dim)
    conv("conv2", z_dim, z_dim, (1, 1, 1))
    conv("decoder.conv1", d0, z_dim, (3, 3, 3))
    res("decoder.middle.0", d0, d0)
    gamma("decoder.middle.1.norm.gamma", d0, True)
    conv("decoder.middle.1.to_qkv", 3 * d0, d0, (1, 1))
    conv("decoder.middle.1.proj", d0, d0, (1, 1))
    res("decoder.middle.2", d0, d0)
    for n, layer in enumerate(layers):
        p = f"decoder.upsamples.{n}"
        if layer[0] == "res":
            res(p, layer[1], layer[2])
        else:
            conv(p + ".resample.1", layer[1] // 2, layer[1], (3, 3))
            if layer[0] == "up3d":
                conv(p + ".time_conv", layer[1] * 2, layer[1], (3, 1, 1))
    gamma("decoder.head.0.gamma", d_out, False)
    conv("decoder.head.2", 3, d_out, (3, 3, 3))
    return W


# ---------------------------------------------------------------------------------------------------------------
# layers
# ---------------------------------------------------------------------------------------------------------------
def causal_conv3d(x, w, b, cache_x=None):
    """CausalConv3d.forward (vae.py:35-44): pad (k-1) frames on the past side only — or use the cached frames — then conv3d."""
    kt, kh, kw = w.shape[2:]
    padding = [kw // 2, kw // 2, kh // 2, kh // 2, kt - 1, 0]
    if cache_x is not None and padding[4] > 0:
        x = torch.cat([cache_x, x], dim=2)
        padding[4] -= cache_x.shape[2]
    return F.conv3d(F.pad(x, padding), w, b)


def rms_norm(x, gamma, channel_dim=1):
    """RMS_norm.forward (vae.py:58-59)."""
    return F.normalize(x, dim=channel_dim) * (x.shape[channel_dim] ** 0.5) * gamma


def _cached_conv(x, w, b, feat_cache, feat_idx):
    """The cache protocol around every 3x3x3 conv (vae.py:203-217, 442-455, 474-487)."""
    idx = feat_idx[0]
    cache_x = x[:, :, -CACHE_T:, :, :].clone()
    if cache_x.shape[2] < 2 and feat_cache[idx] is not None:
        cache_x = torch.cat([feat_cache[idx][:, :, -1, :, :].unsqueeze(2), cache_x], dim=2)
    y = causal_conv3d(x, w, b, feat_cache[idx])
    feat_cache[idx] = cache_x
    feat_idx[0] += 1
    return y


def residual_block(W, p, x, feat_cache, feat_idx):
    """ResidualBlock.forward (vae.py:202-223)."""
    h = causal_conv3d(x, W[p + ".shortcut.weight"], W[p + ".shortcut.bias"]) if (p + ".shortcut.weight") in W else x
    x = F.silu(rms_norm(x, W[p + ".residual.0.gamma"]))
    x = _cached_conv(x, W[p + ".residual.2.weight"], W[p + ".residual.2.bias"], feat_cache, feat_idx)
    x = F.silu(rms_norm(x, W[p + ".residual.3.gamma"]))
    x = _cached_conv(x, W[p + ".residual.6.weight"], W[p + ".residual.6.bias"], feat_cache, feat_idx)
    return x + h


def attention_block(W, p, x):
    """AttentionBlock.forward (vae.py:245-262): per-frame single-head spatial attention."""
    identity = x
    b, c, t, h, w = x.shape
    x = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    x = rms_norm(x, W[p + ".norm.gamma"])
    qkv = F.conv2d(x, W[p + ".to_qkv.weight"], W[p + ".to_qkv.bias"])
    q, k, v = qkv.reshape(b * t, 1, c * 3, -1).permute(0, 1, 3, 2).contiguous().chunk(3, dim=-1)
    x = F.scaled_dot_product_attention(q, k, v)
    x = x.squeeze(1).permute(0, 2, 1).reshape(b * t, c, h, w)
    x = F.conv2d(x, W[p + ".proj.weight"], W[p + ".proj.bias"])
    x = x.reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4)
    return x + identity


def resample_up(W, p, x, mode, feat_cache, feat_idx):
    """Resample.forward, upsample2d / upsample3d (vae.py:107-143)."""
    b, c, t, h, w = x.shape
    if mode == "up3d":
        idx = feat_idx[0]
        if feat_cache[idx] is None:
            feat_cache[idx] = "Rep"                       # first latent frame: no temporal conv, one frame out
            feat_idx[0] += 1
        else:
            cache_x = x[:, :, -CACHE_T:, :, :].clone()
            if cache_x.shape[2] < 2 and not isinstance(feat_cache[idx], str):
                cache_x = torch.cat([feat_cache[idx][:, :, -1, :, :].unsqueeze(2), cache_x], dim=2)
            if cache_x.shape[2] < 2 and isinstance(feat_cache[idx], str):
                cache_x = torch.cat([torch.zeros_like(cache_x), cache_x], dim=2)
            tw, tb = W[p + ".time_conv.weight"], W[p + ".time_conv.bias"]
            x = causal_conv3d(x, tw, tb) if isinstance(feat_cache[idx], str) else causal_conv3d(x, tw, tb, feat_cache[idx])
            feat_cache[idx] = cache_x
            feat_idx[0] += 1
            x = x.reshape(b, 2, c, t, h, w)
            x = torch.stack((x[:, 0], x[:, 1]), 3).reshape(b, c, t * 2, h, w)
    t = x.shape[2]
    x = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    x = F.interpolate(x.float(), scale_factor=(2.0, 2.0), mode="nearest-exact").type_as(x)
    x = F.conv2d(x, W[p + ".resample.1.weight"], W[p + ".resample.1.bias"], padding=1)
    return x.reshape(b, t, c // 2, h * 2, w * 2).permute(0, 2, 1, 3, 4)


def decoder3d(W, x, feat_cache, feat_idx, layers):
    """Decoder3d.forward (vae.py:436-489)."""
    x = _cached_conv(x, W["decoder.conv1.weight"], W["decoder.conv1.bias"], feat_cache, feat_idx)
    x = residual_block(W, "decoder.middle.0", x, feat_cache, feat_idx)
    x = attention_block(W, "decoder.middle.1", x)
    x = residual_block(W, "decoder.middle.2", x, feat_cache, feat_idx)
    for n, layer in enumerate(layers):
        p = f"decoder.upsamples.{n}"
        if layer[0] == "res":
            x = residual_block(W, p, x, feat_cache, feat_idx)
        else:
            x = resample_up(W, p, x, layer[0], feat_cache, feat_idx)
    x = F.silu(rms_norm(x, W["decoder.head.0.gamma"]))
    return _cached_conv(x, W["decoder.head.2.weight"], W["decoder.head.2.bias"], feat_cache, feat_idx)


def count_causal_convs(layers) -> int:
    """count_conv3d(decoder) (vae.py:492-497): every CausalConv3d incl. 1x1x1 shortcuts and time_convs."""
    n = 1 + 2 * 2 + 1                                   # conv1, two middle res blocks, head conv
    for layer in layers:
        if layer[0] == "res":
            n += 2 + (1 if layer[1] != layer[2] else 0)
        elif layer[0] == "up3d":
            n += 1
    return n


def vae_decode(W: Dict[str, torch.Tensor], zs: torch.Tensor, dim=96, z_dim=16) -> torch.Tensor:
    """WanVAE.decode -> WanVAE_.decode (vae.py:931-957, 713-738): zs [16, T, H, W] fp32 -> [1, 3, 1+4(T-1), 8H, 8W] in [-1, 1]."""
    _, layers, _ = decoder_layout(dim, z_dim)
    z = zs.unsqueeze(0)
    mean = torch.tensor(MEAN, dtype=z.dtype, device=z.device)
    inv_std = 1.0 / torch.tensor(STD, dtype=z.dtype, device=z.device)
    z = z / inv_std.view(1, z_dim, 1, 1, 1) + mean.view(1, z_dim, 1, 1, 1)
    x = causal_conv3d(z, W["conv2.weight"], W["conv2.bias"])
    feat_cache: List[Optional[torch.Tensor]] = [None] * count_causal_convs(layers)
    outs = []
    for i in range(z.shape[2]):
        outs.append(decoder3d(W, x[:, :, i : i + 1], feat_cache, [0], layers))
    return torch.cat(outs, dim=2).float().clamp_(-1, 1)


# ---------------------------------------------------------------------------------------------------------------
# Encoder (lightx2v/models/video_encoders/hf/wan/vae.py: Encoder3d :264-376, Resample downsample2d / downsample3d :70-159,
# WanVAE_.encode :684-711).  Pinned by tests/golden/wan_vae_encode_small.safetensors (oracle/gen_golden.py:gen_vae_encode_fixture,
# REAL WanVAE_.encode on CPU in fp32).  Weight keys: `encoder.conv1.*`, `encoder.downsamples.N.*`, `encoder.middle.{0,1,2}.*`,
# `encoder.head.{0,2}.*`, `conv1.*` (the 1x1x1 conv in front of the mu / log_var split).
# ---------------------------------------------------------------------------------------------------------------
def encoder_layout(dim=96, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temperal_downsample=(False, True, True)):
    """Module list of Encoder3d.downsamples (vae.py:291-305): [("res", in, out) | ("down2d"/"down3d", dim)], and the final width."""
    dims = [dim * u for u in [1] + list(dim_mult)]
    layers = []
    out_dim = dims[0]
    for i, (in_dim, out_dim) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(num_res_blocks):
            layers.append(("res", in_dim, out_dim))
            in_dim = out_dim
        if i != len(dim_mult) - 1:
            layers.append(("down3d" if temperal_downsample[i] else "down2d", out_dim))
    return dims[0], layers, out_dim


def synth_vae_encoder_weights(seed=0, dim=96, z_dim=16, device="cpu", dtype=torch.float32) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed + 1000)
    W: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k):
        fan = cin * math.prod(k)
        W[name + ".weight"] = (torch.randn(cout, cin, *k, generator=g) * (1.0 / math.sqrt(fan))).to(dtype).to(device)
        W[name + ".bias"] = (torch.randn(cout, generator=g) * 0.05).to(dtype).to(device)

    def gamma(name, c, images):
        shape = (c, 1, 1) if images else (c, 1, 1, 1)
        W[name] = (1.0 + 0.1 * torch.randn(shape, generator=g)).to(dtype).to(device)

    def res(prefix, cin, cout):
        gamma(prefix + ".residual.0.gamma", cin, False)
        conv(prefix + ".residual.2", cout, cin, (3, 3, 3))
        gamma(prefix + ".residual.3.gamma", cout, False)
        conv(prefix + ".residual.6", cout, cout, (3, 3, 3))
        if cin != cout:
            conv(prefix + ".shortcut", cout, cin, (1, 1, 1))

    d0, layers, d_out = encoder_layout(dim)
    conv("encoder.conv1", d0, 3, (3, 3, 3))
    for n, layer in enumerate(layers):
        p = f"encoder.downsamples.{n}"
        if layer[0] == "res":
            res(p, layer[1], layer[2])
        else:
            conv(p + ".resample.1", layer[1], layer[1], (3, 3))
            if layer[0] == "down3d":
                conv(p + ".time_conv", layer[1], layer[1], (3, 1, 1))
    res("encoder.middle.0", d_out, d_out)
    gamma("encoder.middle.1.norm.gamma", d_out, True)
    conv("encoder.middle.1.to_qkv", 3 * d_out, d_out, (1, 1))
    conv("encoder.middle.1.proj", d_out, d_out, (1, 1))
    res("encoder.middle.2", d_out, d_out)
    gamma("encoder.head.0.gamma", d_out, False)
    conv("encoder.head.2", 2 * z_dim, d_out, (3, 3, 3))
    conv("conv1", 2 * z_dim, 2 * z_dim, (1, 1, 1))
    return W


def resample_down(W, p, x, mode, feat_cache, feat_idx):
    """Resample.forward, downsample2d / downsample3d (vae.py:139-159): ZeroPad2d((0, 1, 0, 1)) + Conv2d(3, stride 2) per frame; then
    (down3d) the first chunk passes through and is cached, later chunks run time_conv (3x1x1, stride 2 in time, no padding) over
    [last cached frame ; chunk]."""
    b, c, t, h, w = x.shape
    x = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    x = F.conv2d(F.pad(x, (0, 1, 0, 1)), W[p + ".resample.1.weight"], W[p + ".resample.1.bias"], stride=(2, 2))
    x = x.reshape(b, t, c, x.shape[-2], x.shape[-1]).permute(0, 2, 1, 3, 4)
    if mode == "down3d":
        idx = feat_idx[0]
        if feat_cache[idx] is None:
            feat_cache[idx] = x.clone()
            feat_idx[0] += 1
        else:
            cache_x = x[:, :, -1:, :, :].clone()
            x = F.conv3d(torch.cat([feat_cache[idx][:, :, -1:, :, :], x], 2), W[p + ".time_conv.weight"], W[p + ".time_conv.bias"], stride=(2, 1, 1))
            feat_cache[idx] = cache_x
            feat_idx[0] += 1
    return x


def encoder3d(W, x, feat_cache, feat_idx, layers):
    """Encoder3d.forward (vae.py:323-376)."""
    x = _cached_conv(x, W["encoder.conv1.weight"], W["encoder.conv1.bias"], feat_cache, feat_idx)
    for n, layer in enumerate(layers):
        p = f"encoder.downsamples.{n}"
        if layer[0] == "res":
            x = residual_block(W, p, x, feat_cache, feat_idx)
        else:
            x = resample_down(W, p, x, layer[0], feat_cache, feat_idx)
    x = residual_block(W, "encoder.middle.0", x, feat_cache, feat_idx)
    x = attention_block(W, "encoder.middle.1", x)
    x = residual_block(W, "encoder.middle.2", x, feat_cache, feat_idx)
    x = F.silu(rms_norm(x, W["encoder.head.0.gamma"]))
    return _cached_conv(x, W["encoder.head.2.weight"], W["encoder.head.2.bias"], feat_cache, feat_idx)


def vae_encode(W: Dict[str, torch.Tensor], video: torch.Tensor, dim=96, z_dim=16) -> torch.Tensor:
    """WanVAE_.encode (vae.py:684-711): video [3, T, H, W] in [-1, 1], T = 1 + 4k -> normalised mu [16, 1 + k, H/8, W/8]; frames are
    fed as 1, 4, 4, ... with the per-convolution caches carried across the chunks."""
    _, layers, _ = encoder_layout(dim)
    x = video.unsqueeze(0)
    n_cache = 64                                        # upper bound on the CausalConv3d count (count_conv3d, vae.py:492-497)
    feat_cache: List[Optional[torch.Tensor]] = [None] * n_cache
    t = x.shape[2]
    outs = []
    for i in range(1 + (t - 1) // 4):
        chunk = x[:, :, :1] if i == 0 else x[:, :, 1 + 4 * (i - 1): 1 + 4 * i]
        outs.append(encoder3d(W, chunk, feat_cache, [0], layers))
    out = torch.cat(outs, 2)
    mu, _ = causal_conv3d(out, W["conv1.weight"], W["conv1.bias"]).chunk(2, dim=1)
    mean = torch.tensor(MEAN, dtype=mu.dtype, device=mu.device).view(1, z_dim, 1, 1, 1)
    inv_std = (1.0 / torch.tensor(STD, dtype=mu.dtype, device=mu.device)).view(1, z_dim, 1, 1, 1)
    return ((mu - mean) * inv_std)[0]
