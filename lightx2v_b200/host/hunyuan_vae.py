"""HunyuanVideo causal 3-D VAE decoder on the sm_100a implicit-GEMM convolution (csrc/conv3d.cu) + GroupNorm kernels (csrc/gn.cu),
with the reference's interface `VideoEncoderKLCausal3DModel.decode(latents, generator, config) -> images [1, 3, 1+4(T-1), 8H, 8W]
fp32 on the CPU in [0, 1]` (lightx2v/models/video_encoders/hf/autoencoder_kl_causal_3d/model.py:33-44).

Structure followed (same directory): AutoencoderKLCausal3D._decode / temporal_tiled_decode / spatial_tiled_decode / blend_*
(autoencoder_kl_causal_3d.py:287-302, 487-518, 405-451, 363-379) -> DecoderCausal3D.forward (vae.py:221-283) -> UNetMidBlockCausal3D
(unet_causal_3d_blocks.py:575-588) / UpDecoderBlockCausal3D (:750-758) / ResnetBlockCausal3D (:364-419) / UpsampleCausal3D (:146-200)
/ CausalConv3d (:65-91).  The tiling (overlapping tiles decoded independently, linear blends) is part of the reference's numerics
- GroupNorm statistics are per tile - so it is reproduced tile for tile, not removed.

B200 mapping, per tile, activations channels-last bf16 [T, H, W, C]:
  * GroupNorm -> SiLU -> replicate pad is ONE write: gn_stats (1 read) + gn_apply_pad (1 read, 1 write of the padded tensor); the
    convolution then reads the padded tensor through TMA with non-negative tap offsets (no padding logic, no F.pad copy);
  * every 3x3x3 convolution is the tcgen05 implicit GEMM; residual / shortcut adds live in its epilogue;
  * nearest-upsample + conv is phase-decomposed on the ORIGINAL resolution: (2,2,2) -> 8 phase convolutions of 2x2x2 taps
    (pre-summed weights; 3.4x fewer FLOPs, the 8x larger upsampled tensor is never materialised), (1,2,2) -> 4 phases of 3x2x2.
    Temporal phases honour the first-frame rule of UpsampleCausal3D (frame 0 is not duplicated) and the causal replicate pad:
    out[2m] = w0 f[m-1] + (w1+w2) f[m],  out[2m+1] = (w0+w1) f[m] + w2 f[m+1]   (f[-1] = f[0]);
  * the single-head 512-wide mid-block attention (< 1 % of the FLOPs) uses the bf16 GEMM kernel for QKV / out projections and torch
    SDPA per frame for the frame-causal softmax (library call; frame i attends frames <= i, so no mask tensor is needed).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .. import lib
from .wan_vae import _Conv

HUNYUAN_VAE_CONFIG = {"block_out_channels": (128, 256, 512, 512), "layers_per_block": 2, "latent_channels": 16, "sample_size": 256,
                      "sample_tsize": 64, "scaling_factor": 0.476986, "tile_overlap_factor": 0.25}

PAD = (2, 1, 1)     # replicate border carried by every padded activation: 2 frames in front, 1 pixel around


class _PConv(_Conv):
    """Convolution over a pre-padded input (taps = non-negative offsets into it)."""

    def __call__(self, xp: torch.Tensor, out: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, extent=None):
        if out is None:
            T, H, W = extent if extent is not None else (xp.shape[0] - PAD[0], xp.shape[1] - 2 * PAD[1], xp.shape[2] - 2 * PAD[2])
            out = torch.empty((T, H, W, self.cout), dtype=torch.bfloat16, device=xp.device)
        return lib.conv3d_cl_padded(xp, self.weight, self.bias, out, self.taps, residual=residual)


def _conv333(w: torch.Tensor, b: torch.Tensor, device, cin_pad=None, cout_pad=None) -> _PConv:
    cout, cin = w.shape[:2]
    taps = [(it, ih, iw) for it in range(3) for ih in range(3) for iw in range(3)]
    wm = w.float().cpu().permute(0, 2, 3, 4, 1).reshape(cout, 27, cin)
    return _PConv(wm, b, device, cin_pad=cin_pad, cout_pad=cout_pad, taps=taps)


class _UpsampleConv3d:
    """UpsampleCausal3D (nearest, factor (ft, 2, 2), first frame not duplicated) + CausalConv3d 3x3x3, phase-decomposed."""

    def __init__(self, w: torch.Tensor, b: torch.Tensor, device, ft: int):
        w = w.float().cpu()
        self.cout, self.ft = w.shape[0], ft
        # (offset into the padded source, kernel indices summed) per output phase; padded coords: f[m] <-> m + 2, h <-> h + 1
        sp = {0: [(0, [0]), (1, [1, 2])], 1: [(1, [0, 1]), (2, [2])]}
        tm = {0: [(1, [0]), (2, [1, 2])], 1: [(2, [0, 1]), (3, [2])]} if ft == 2 else {0: [(0, [0]), (1, [1]), (2, [2])]}
        self.phases = []
        for pt_, trows in tm.items():
            for py in (0, 1):
                for px in (0, 1):
                    taps, mats = [], []
                    for dt, kts in trows:
                        for dh, khs in sp[py]:
                            for dw, kws in sp[px]:
                                taps.append((dt, dh, dw))
                                mats.append(sum(w[:, :, kt, kh, kw] for kt in kts for kh in khs for kw in kws))
                    self.phases.append((pt_, py, px, _PConv(torch.stack(mats, dim=1), b, device, taps=taps)))

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        T, H, W, C = x.shape
        xp = lib.gn_apply_pad_cl(x, None, None, None, pad=PAD)
        Tout = 2 * T - 1 if self.ft == 2 else T
        y = torch.empty((Tout, 2 * H, 2 * W, self.cout), dtype=torch.bfloat16, device=x.device)
        for pt_, py, px, conv in self.phases:
            view = y[pt_::2, py::2, px::2] if self.ft == 2 else y[:, py::2, px::2]
            if view.shape[0] > 0:
                conv(xp, out=view)
        return y


class HunyuanVAEDecoderB200:
    """DecoderCausal3D + post_quant_conv on one tile."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", config=None):
        self.cfg = dict(HUNYUAN_VAE_CONFIG, **(config or {}))
        W = state_dict
        dev = self.device = torch.device(device)
        chans = list(reversed(self.cfg["block_out_channels"]))
        n_blocks = len(chans)
        self.zc = self.cfg["latent_channels"]

        def f32(name):
            return W[name].float().reshape(-1).to(dev).contiguous()

        def norm(p):
            return (f32(p + ".weight"), f32(p + ".bias"))

        def res(p):
            d = {"n1": norm(p + ".norm1"), "c1": _conv333(W[p + ".conv1.conv.weight"], W[p + ".conv1.conv.bias"], dev),
                 "n2": norm(p + ".norm2"), "c2": _conv333(W[p + ".conv2.conv.weight"], W[p + ".conv2.conv.bias"], dev), "sc": None}
            if (p + ".conv_shortcut.conv.weight") in W:
                d["sc"] = _Conv(W[p + ".conv_shortcut.conv.weight"], W[p + ".conv_shortcut.conv.bias"], dev)
            return d

        # latent [16] is zero-padded to 64 channels so that every conv of the stack can use the same K-chunk geometry
        self.post_quant = _Conv(W["post_quant_conv.weight"], W["post_quant_conv.bias"], dev, cin_pad=64, cout_pad=64)
        self.conv_in = _conv333(W["decoder.conv_in.conv.weight"], W["decoder.conv_in.conv.bias"], dev, cin_pad=64)
        self.mid0 = res("decoder.mid_block.resnets.0")
        a = "decoder.mid_block.attentions.0"
        bf = lambda n: W[n].to(torch.bfloat16).to(dev).contiguous()     # noqa: E731
        self.attn = {"norm": norm(a + ".group_norm"),
                     "wqkv": torch.cat([bf(f"{a}.to_{n}.weight") for n in "qkv"], 0).contiguous(),
                     "bqkv": torch.cat([bf(f"{a}.to_{n}.bias") for n in "qkv"], 0).contiguous(),
                     "wo": bf(a + ".to_out.0.weight"), "bo": bf(a + ".to_out.0.bias")}
        self.mid1 = res("decoder.mid_block.resnets.1")
        self.layers: List[Tuple[str, object]] = []
        for i in range(n_blocks):
            for j in range(self.cfg["layers_per_block"] + 1):
                self.layers.append(("res", res(f"decoder.up_blocks.{i}.resnets.{j}")))
            sp = i < 3
            tm = i >= n_blocks - 1 - 2 and i != n_blocks - 1
            if sp or tm:
                assert sp, "time-only upsampling does not occur in the 4-block decoder"
                p = f"decoder.up_blocks.{i}.upsamplers.0.conv.conv"
                self.layers.append(("up", _UpsampleConv3d(W[p + ".weight"], W[p + ".bias"], dev, 2 if tm else 1)))
        self.norm_out = norm("decoder.conv_norm_out")
        self.conv_out = _conv333(W["decoder.conv_out.conv.weight"], W["decoder.conv_out.conv.bias"], dev, cout_pad=16)
        self._sums = torch.empty(int(lib.load().b200_gn_stats_workspace_doubles()), dtype=torch.float64, device=dev)
        self._zero = torch.zeros(self.zc, dtype=torch.float32, device=dev)
        self._scale = torch.full((self.zc,), float(self.cfg["scaling_factor"]), dtype=torch.float32, device=dev)

    # ------------------------------------------------------------------ blocks
    def _gn_pad(self, x, nb, silu=True, pad=PAD):
        return lib.gn_apply_pad_cl(x, lib.gn_stats_cl(x, self._sums), nb[0], nb[1], eps=1e-6, pad=pad, silu=silu)

    def _res(self, d, x):
        h = x if d["sc"] is None else d["sc"](x)
        a = d["c1"](self._gn_pad(x, d["n1"]))
        return d["c2"](self._gn_pad(a, d["n2"]), residual=h)

    def _attention(self, x):
        T, H, W, C = x.shape
        hw = H * W
        n = self._gn_pad(x, self.attn["norm"], silu=False, pad=(0, 0, 0)).view(T * hw, C)
        qkv = lib.gemm_bf16(n, self.attn["wqkv"], self.attn["bqkv"])
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
        o = torch.empty((T * hw, C), dtype=torch.bfloat16, device=x.device)
        for i in range(T):          # prepare_causal_attention_mask: frame i sees frames <= i
            L = (i + 1) * hw
            o[i * hw:L] = F.scaled_dot_product_attention(q[None, None, i * hw:L], k[None, None, :L], v[None, None, :L])[0, 0]
        out = x.reshape(T * hw, C).clone()
        lib.gemm_bf16(o, self.attn["wo"], self.attn["bo"], out=out, epilogue=lib.EPI_RESIDUAL)
        return out.view(T, H, W, C)

    @torch.no_grad()
    def decode_tile(self, z: torch.Tensor) -> torch.Tensor:
        """z [16, T, h, w] fp32 (already divided by nothing: raw latents) -> [3, T', 8h, 8w] fp32 (pre-clamp decoder output)."""
        z = z.to(self.device, torch.float32).contiguous()
        x = lib.latent_to_cl(z, self._zero, self._scale, cp=64)                  # z / scaling_factor, channels-last, 64-padded
        x = self.post_quant(x)
        x = self.conv_in(lib.gn_apply_pad_cl(x, None, None, None, pad=PAD))
        x = self._res(self.mid0, x)
        x = self._attention(x)
        x = self._res(self.mid1, x)
        for kind, layer in self.layers:
            x = self._res(layer, x) if kind == "res" else layer(x)
        x = self.conv_out(self._gn_pad(x, self.norm_out))
        return lib.cl_to_video(x)


def _blend(a: torch.Tensor, b: torch.Tensor, extent: int, dim: int) -> torch.Tensor:
    """blend_v / blend_h / blend_t (autoencoder_kl_causal_3d.py:363-379), vectorised; in place on b like the reference."""
    extent = min(a.shape[dim], b.shape[dim], extent)
    if extent <= 0:
        return b
    shape = [1] * b.dim()
    shape[dim] = extent
    r = (torch.arange(extent, device=b.device, dtype=torch.float32) / extent).view(shape)
    bs = b.narrow(dim, 0, extent)
    bs.copy_(a.narrow(dim, a.shape[dim] - extent, extent) * (1 - r) + bs * r)
    return b


class HunyuanVAEB200:
    """Drop-in for VideoEncoderKLCausal3DModel (model.py:6-44): `.decode(latents, generator, config)`; tiling always enabled."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", config=None):
        self.decoder = HunyuanVAEDecoderB200(state_dict, device, config)
        self.cfg = self.decoder.cfg
        self.device = self.decoder.device
        c = self.cfg
        self.tile_sample = c["sample_size"]
        self.tile_latent = int(c["sample_size"] / (2 ** (len(c["block_out_channels"]) - 1)))
        self.tile_tsample = c["sample_tsize"]
        self.tile_tlatent = c["sample_tsize"] // 4
        self.overlap = c["tile_overlap_factor"]
        self._dist = None
        self._tile_counter = 0

    # ------------------------------------------------------------------ tile scheduling (single GPU: every tile is local)
    def _decode_tile(self, z: torch.Tensor) -> torch.Tensor:
        """Decode one tile here, or - in a tile-parallel decode (decode_dist) - hand out the tile that the owning rank decoded.
        Tiles are independent units of work (the reference decodes them one after another), so sharding them needs no collective on the
        compute path.  decode_dist runs the tiling traversal twice: a PLAN pass that only records the tile inputs in traversal order,
        then (after every rank has decoded its own tiles back to back and the finished tiles have been exchanged) the real pass, which
        pops the finished tiles in the same order and performs the blends."""
        if self._dist is None:
            return self.decoder.decode_tile(z)
        mode = self._dist[3]
        if mode == "plan":
            self._plan.append(z)
            _, t, h, w = z.shape
            return torch.empty((3, 1 + 4 * (t - 1), 0, 0), dtype=torch.float32, device=self.device)      # shape carrier only (no blends in the plan pass)
        out = self._done[self._tile_counter]
        self._done[self._tile_counter] = None
        self._tile_counter += 1
        return out

    def _spatial_tiled(self, z: torch.Tensor) -> torch.Tensor:
        step = int(self.tile_latent * (1 - self.overlap))
        extent = int(self.tile_sample * self.overlap)
        limit = self.tile_sample - extent
        rows = []
        for i in range(0, z.shape[-2], step):
            rows.append([self._decode_tile(z[:, :, i:i + self.tile_latent, j:j + self.tile_latent]) for j in range(0, z.shape[-1], step)])
        if self._dist is not None and self._dist[3] == "plan":
            return rows[0][0]
        out_rows = []
        for i, row in enumerate(rows):
            out = []
            for j, tile in enumerate(row):
                if i > 0:
                    tile = _blend(rows[i - 1][j], tile, extent, -2)
                if j > 0:
                    tile = _blend(row[j - 1], tile, extent, -1)
                out.append(tile[:, :, :limit, :limit])
            out_rows.append(torch.cat(out, dim=-1))
        return torch.cat(out_rows, dim=-2)

    def _temporal_tiled(self, z: torch.Tensor) -> torch.Tensor:
        step = int(self.tile_tlatent * (1 - self.overlap))
        extent = int(self.tile_tsample * self.overlap)
        limit = self.tile_tsample - extent
        row = []
        for i in range(0, z.shape[1], step):
            tile = z[:, i:i + self.tile_tlatent + 1]
            if tile.shape[-1] > self.tile_latent or tile.shape[-2] > self.tile_latent:
                dec = self._spatial_tiled(tile)
            else:
                dec = self._decode_tile(tile)
            if i > 0:
                dec = dec[:, 1:]
            row.append(dec)
        if self._dist is not None and self._dist[3] == "plan":
            return row[0]
        out = []
        for i, tile in enumerate(row):
            if i > 0:
                tile = _blend(row[i - 1], tile, extent, 1)
                out.append(tile[:, :limit])
            else:
                out.append(tile[:, :limit + 1])
        return torch.cat(out, dim=1)

    def _traverse(self, z: torch.Tensor) -> torch.Tensor:
        if z.shape[1] > self.tile_tlatent:
            return self._temporal_tiled(z)
        if z.shape[-1] > self.tile_latent or z.shape[-2] > self.tile_latent:
            return self._spatial_tiled(z)
        return self._decode_tile(z)

    @torch.no_grad()
    def decode_device(self, latents: torch.Tensor) -> torch.Tensor:
        """latents [1, 16, T, H, W] -> [1, 3, T', 8H, 8W] fp32 in [0, 1], still on the GPU."""
        z = latents.to(self.device, torch.float32)[0]
        return self._traverse(z).mul_(0.5).add_(0.5).clamp_(0, 1).unsqueeze(0)

    @torch.no_grad()
    def decode_dist(self, latents: torch.Tensor, group=None, to_cpu: bool = True) -> torch.Tensor:
        """Tile-parallel decode over the ranks of `group`: tile k of the reference's (temporal, row, column) tile order is decoded by
        rank k mod P; every rank decodes ITS tiles back to back (no communication in between), then the finished tiles are exchanged
        (one broadcast per tile, all enqueued together) and every rank performs the same blends and returns the same video -
        bit-identical to `decode`, since each tile is computed by the same kernels on the same inputs.  (Round 1 interleaved one
        broadcast after every tile, which serialised the ranks behind each other: 33 MPix/s on 8 GPUs against 29 on one.)  The
        reference has no parallel Hunyuan VAE (its Wan VAE has `parallel_vae`, hf/wan/vae.py:883-929); with 84 independent tiles at
        720p x 129f this is the natural B200 sharding."""
        import torch.distributed as dist

        world, rank = dist.get_world_size(group), dist.get_rank(group)
        z = latents.to(self.device, torch.float32)[0]
        self._plan, self._done, self._tile_counter = [], [], 0
        try:
            self._dist = (world, rank, group, "plan")
            self._traverse(z)                                            # records the tile inputs in traversal order
            tiles = self._plan
            done = [None] * len(tiles)
            for k, zt in enumerate(tiles):                               # my tiles, back to back
                if k % world == rank:
                    done[k] = self.decoder.decode_tile(zt)
            for k, zt in enumerate(tiles):                               # hand-over of the finished tiles
                if done[k] is None:
                    _, t, h, w = zt.shape
                    done[k] = torch.empty((3, 1 + 4 * (t - 1), 8 * h, 8 * w), dtype=torch.float32, device=self.device)
                owner = k % world
                dist.broadcast(done[k], src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
            self._done, self._tile_counter = done, 0
            self._dist = (world, rank, group, "blend")
            img = self._traverse(z).mul_(0.5).add_(0.5).clamp_(0, 1).unsqueeze(0)
        finally:
            self._dist, self._plan, self._done = None, [], []
        return img.cpu().float() if to_cpu else img

    @torch.no_grad()
    def decode(self, latents: torch.Tensor, generator=None, config=None) -> torch.Tensor:
        return self.decode_device(latents).cpu().float()
