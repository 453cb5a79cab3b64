"""ORACLE (test infrastructure, NOT product code): plain-torch restatement of the HunyuanVideo causal 3-D VAE DECODE path as
LightX2V runs it — VideoEncoderKLCausal3DModel.decode (lightx2v/models/video_encoders/hf/autoencoder_kl_causal_3d/model.py:33-44)
-> AutoencoderKLCausal3D.temporal_tiled_decode / spatial_tiled_decode / blend_* (autoencoder_kl_causal_3d.py:487-518, 405-451,
363-379) -> DecoderCausal3D.forward (vae.py:221-283) -> UNetMidBlockCausal3D / UpDecoderBlockCausal3D / ResnetBlockCausal3D /
UpsampleCausal3D / CausalConv3d (unet_causal_3d_blocks.py:575-588, 750-758, 364-419, 146-200, 65-91).

Weights: flat state_dict with the reference module's own key names (`decoder.conv_in.conv.weight`, `post_quant_conv.weight` ...).

Third-party arithmetic on this path that is NOT under /root/reference: the single-head mid-block `Attention` comes from
`diffusers.models.attention_processor` (HunyuanVideo pins diffusers 0.31; the package is absent from this image).  Its
published algorithm for `_from_deprecated_attn_block=True, residual_connection=True` is restated in `mid_attention` below:
GroupNorm over [B, C, S] -> to_q/to_k/to_v Linear -> softmax(q k^T / sqrt(C) + mask) v -> to_out[0] Linear -> (+ input) /
rescale_output_factor.  Pinned by tests/golden/hunyuan_vae_decode_small.safetensors, produced by the REAL reference classes
(oracle/gen_golden.py, with a diffusers shim that supplies only that Attention layer and the config mixins)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


def causal_conv3d(W, name, x):
    """CausalConv3d.forward — unet_causal_3d_blocks.py:65-91: replicate pad (k//2 each side of W and H, k-1 frames in front)."""
    w = W[name + ".conv.weight"]
    k = w.shape[-1]
    if k > 1:
        x = F.pad(x, (k // 2, k // 2, k // 2, k // 2, k - 1, 0), mode="replicate")
    return F.conv3d(x, w, W.get(name + ".conv.bias"))


def group_norm(W, name, x, groups=32, eps=1e-6):
    return F.group_norm(x, groups, W[name + ".weight"], W[name + ".bias"], eps)


def resnet(W, p, x):
    """ResnetBlockCausal3D.forward — unet_causal_3d_blocks.py:364-419 (temb None, output_scale_factor 1, dropout 0)."""
    h = F.silu(group_norm(W, p + ".norm1", x))
    h = causal_conv3d(W, p + ".conv1", h)
    h = F.silu(group_norm(W, p + ".norm2", h))
    h = causal_conv3d(W, p + ".conv2", h)
    if (p + ".conv_shortcut.conv.weight") in W:
        x = causal_conv3d(W, p + ".conv_shortcut", x)
    return x + h


def upsample(W, p, x, factor):
    """UpsampleCausal3D.forward — :146-200: the first frame is upsampled only spatially, the others by `factor` (nearest)."""
    B, C, T, H, Wd = x.shape
    first, other = x.split((1, T - 1), dim=2)
    first = F.interpolate(first.view(B, C, H, Wd), scale_factor=factor[1:], mode="nearest").unsqueeze(2)
    if T > 1:
        other = F.interpolate(other, scale_factor=factor, mode="nearest")
        x = torch.cat((first, other), dim=2)
    else:
        x = first
    return causal_conv3d(W, p + ".conv", x)


def causal_frame_mask(n_frame, n_hw, dtype, device):
    """prepare_causal_attention_mask — :44-62: a token of frame i sees every token of frames <= i."""
    f = torch.arange(n_frame * n_hw, device=device) // n_hw
    m = torch.zeros(n_frame * n_hw, n_frame * n_hw, dtype=dtype, device=device)
    m.masked_fill_(f[None, :] > f[:, None], float("-inf"))
    return m


def mid_attention(W, p, x):
    """UNetMidBlockCausal3D.forward attention branch — :577-586, with diffusers' deprecated-attn-block Attention (see header)."""
    B, C, T, H, Wd = x.shape
    s = x.permute(0, 2, 3, 4, 1).reshape(B, T * H * Wd, C)
    res = s
    n = F.group_norm(s.transpose(1, 2), 32, W[p + ".group_norm.weight"], W[p + ".group_norm.bias"], 1e-6).transpose(1, 2)
    q = F.linear(n, W[p + ".to_q.weight"], W[p + ".to_q.bias"])
    k = F.linear(n, W[p + ".to_k.weight"], W[p + ".to_k.bias"])
    v = F.linear(n, W[p + ".to_v.weight"], W[p + ".to_v.bias"])
    a = torch.softmax((q @ k.transpose(1, 2)) * (C ** -0.5) + causal_frame_mask(T, H * Wd, q.dtype, q.device), dim=-1) @ v
    o = F.linear(a, W[p + ".to_out.0.weight"], W[p + ".to_out.0.bias"]) + res
    return o.reshape(B, T, H, Wd, C).permute(0, 4, 1, 2, 3)


def decoder_forward(W, z, cfg):
    """DecoderCausal3D.forward — vae.py:221-283 (structure built in __init__ :140-219)."""
    chans = list(reversed(cfg["block_out_channels"]))
    n_blocks = len(chans)
    n_sp = {8: 3, 4: 2}[cfg.get("spatial_compression_ratio", 8)]
    n_t = 2                                                        # time_compression_ratio == 4
    x = causal_conv3d(W, "decoder.conv_in", z)
    x = resnet(W, "decoder.mid_block.resnets.0", x)
    x = mid_attention(W, "decoder.mid_block.attentions.0", x)
    x = resnet(W, "decoder.mid_block.resnets.1", x)
    for i in range(n_blocks):
        for j in range(cfg["layers_per_block"] + 1):
            x = resnet(W, f"decoder.up_blocks.{i}.resnets.{j}", x)
        sp = i < n_sp
        tm = i >= n_blocks - 1 - n_t and i != n_blocks - 1
        if sp or tm:
            x = upsample(W, f"decoder.up_blocks.{i}.upsamplers.0", x, (2 if tm else 1, 2 if sp else 1, 2 if sp else 1))
    x = F.silu(group_norm(W, "decoder.conv_norm_out", x))
    return causal_conv3d(W, "decoder.conv_out", x)


def tile_decode(W, z, cfg):
    z = F.conv3d(z, W["post_quant_conv.weight"], W["post_quant_conv.bias"])
    return decoder_forward(W, z, cfg)


def tile_params(cfg):
    """AutoencoderKLCausal3D.__init__ — autoencoder_kl_causal_3d.py:119-127."""
    sample = cfg["sample_size"]
    lat = int(sample / (2 ** (len(cfg["block_out_channels"]) - 1)))
    return {"sample": sample, "lat": lat, "tsample": cfg["sample_tsize"], "tlat": cfg["sample_tsize"] // 4, "overlap": 0.25}


def blend(a, b, extent, dim):
    """blend_v / blend_h / blend_t — :363-379 (in place on b, linear ramp over `extent` positions)."""
    extent = min(a.shape[dim], b.shape[dim], extent)
    for i in range(extent):
        bi = b.select(dim, i)
        bi.copy_(a.select(dim, a.shape[dim] - extent + i) * (1 - i / extent) + bi * (i / extent))
    return b


def spatial_tiled_decode(W, z, cfg):
    """:405-451."""
    tp = tile_params(cfg)
    overlap = int(tp["lat"] * (1 - tp["overlap"]))
    extent = int(tp["sample"] * tp["overlap"])
    limit = tp["sample"] - extent
    rows = []
    for i in range(0, z.shape[-2], overlap):
        rows.append([tile_decode(W, z[:, :, :, i:i + tp["lat"], j:j + tp["lat"]], cfg) for j in range(0, z.shape[-1], overlap)])
    out_rows = []
    for i, row in enumerate(rows):
        out = []
        for j, tile in enumerate(row):
            if i > 0:
                tile = blend(rows[i - 1][j], tile, extent, -2)
            if j > 0:
                tile = blend(row[j - 1], tile, extent, -1)
            out.append(tile[:, :, :, :limit, :limit])
        out_rows.append(torch.cat(out, dim=-1))
    return torch.cat(out_rows, dim=-2)


def temporal_tiled_decode(W, z, cfg):
    """:487-518."""
    tp = tile_params(cfg)
    overlap = int(tp["tlat"] * (1 - tp["overlap"]))
    extent = int(tp["tsample"] * tp["overlap"])
    limit = tp["tsample"] - extent
    row = []
    for i in range(0, z.shape[2], overlap):
        tile = z[:, :, i:i + tp["tlat"] + 1]
        if tile.shape[-1] > tp["lat"] or tile.shape[-2] > tp["lat"]:
            dec = spatial_tiled_decode(W, tile, cfg)
        else:
            dec = tile_decode(W, tile, cfg)
        if i > 0:
            dec = dec[:, :, 1:]
        row.append(dec)
    out = []
    for i, tile in enumerate(row):
        if i > 0:
            tile = blend(row[i - 1], tile, extent, 2)
            out.append(tile[:, :, :limit])
        else:
            out.append(tile[:, :, :limit + 1])
    return torch.cat(out, dim=2)


def decode(W: Dict[str, torch.Tensor], latents: torch.Tensor, cfg) -> torch.Tensor:
    """VideoEncoderKLCausal3DModel.decode — model.py:33-44 with enable_tiling(); `_decode` dispatch :287-302."""
    tp = tile_params(cfg)
    z = latents / cfg["scaling_factor"]
    if z.shape[2] > tp["tlat"]:
        img = temporal_tiled_decode(W, z, cfg)
    elif z.shape[-1] > tp["lat"] or z.shape[-2] > tp["lat"]:
        img = spatial_tiled_decode(W, z, cfg)
    else:
        img = tile_decode(W, z, cfg)
    return (img / 2 + 0.5).clamp(0, 1).float()


HUNYUAN_VAE_CFG = {"block_out_channels": (128, 256, 512, 512), "layers_per_block": 2, "latent_channels": 16, "sample_size": 256,
                   "sample_tsize": 64, "scaling_factor": 0.476986}     # hunyuan-video-t2v-720p/vae/config.json (public checkpoint)


def synth_vae_weights(cfg, seed=42) -> Dict[str, torch.Tensor]:
    """Seeded random weights with the reference's key names and shapes (fan-in scaled so activations stay O(1))."""
    g = torch.Generator().manual_seed(seed)
    W: Dict[str, torch.Tensor] = {}
    chans = list(reversed(cfg["block_out_channels"]))
    zc = cfg["latent_channels"]

    def conv(name, cout, cin, k, inner=True):
        key = name + (".conv" if inner else "")
        W[key + ".weight"] = torch.randn(cout, cin, k, k, k, generator=g) * (1.0 / (cin * k ** 3)) ** 0.5
        W[key + ".bias"] = torch.randn(cout, generator=g) * 0.02

    def norm(name, c):
        W[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        W[name + ".bias"] = 0.05 * torch.randn(c, generator=g)

    def res(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cout, cin, 3)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".conv_shortcut", cout, cin, 1)

    conv("post_quant_conv", zc, zc, 1, inner=False)
    conv("decoder.conv_in", chans[0], zc, 3)
    res("decoder.mid_block.resnets.0", chans[0], chans[0])
    a = "decoder.mid_block.attentions.0"
    norm(a + ".group_norm", chans[0])
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        W[f"{a}.{n}.weight"] = torch.randn(chans[0], chans[0], generator=g) * chans[0] ** -0.5
        W[f"{a}.{n}.bias"] = torch.randn(chans[0], generator=g) * 0.02
    res("decoder.mid_block.resnets.1", chans[0], chans[0])
    prev = chans[0]
    for i, c in enumerate(chans):
        for j in range(cfg["layers_per_block"] + 1):
            res(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else c, c)
        if i != len(chans) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c, 3)
        prev = c
    norm("decoder.conv_norm_out", chans[-1])
    conv("decoder.conv_out", 3, chans[-1], 3)
    return W
