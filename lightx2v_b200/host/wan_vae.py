"""Wan 3-D causal VAE decoder on the sm_100a implicit-GEMM convolution (csrc/conv3d.cu), with the reference's interface
`WanVAE.decode(zs, generator, config) -> images [1, 3, 1+4(T-1), 8H, 8W] fp32 in [-1, 1]`
(lightx2v/models/video_encoders/hf/wan/vae.py:931-957; model structure Decoder3d :377-489, _video_vae :762-786).

The reference decodes ONE latent frame per iteration and threads a two-frame cache through every causal convolution
(vae.py:713-738, 203-217).  That protocol is algebraically a causal, zero-padded convolution over the whole frame sequence, with
two special cases that are reproduced here:
  * `upsample3d`: the first latent frame skips the temporal convolution and yields one frame; every later frame t yields two
    frames from time_conv(f[t-2], f[t-1], f[t]) where the history starts at frame 1 (frames < 1 read as zero) (vae.py:107-138);
  * every other conv simply sees zeros before frame 0 (vae.py:35-44).
Processing CHUNKS of latent frames (default 3, `chunk_frames`) instead of one turns 21 x ~35 small cuDNN launches + torch.cat cache
copies into 7 x ~60 large tensor-core launches, while every temporal convolution carries the last two frames of its input from chunk to
chunk - the reference's feat_cache protocol with a chunk size > 1 - so peak memory is bounded by the chunk (~10 GB at 720p) instead of
the whole 81-frame sequence (67 GB in round 1).  `chunk_frames=None` decodes the whole sequence in one piece (bit-identical result).
Further fusions: the 3x3 conv on the nearest-2x-upsampled image is evaluated as four 2x2 phase convolutions on the ORIGINAL
resolution (pre-summed weights): 2.25x fewer FLOPs and the 4x larger upsampled tensor is never materialised; residual adds,
bias and the final clamp live in the conv epilogue.  Activations are channels-last bf16 (the reference computes in fp32 / TF32).
The single-head 384-wide spatial attention of the middle block (1 % of the FLOPs) goes through torch SDPA (library call).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .. import lib

MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]
STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]   # vae.py:804-839


def _pad_to(n: int, m: int) -> int:
    return (n + m - 1) // m * m


class _Conv:
    """A prepared convolution: bf16 weight matrix [cout_p, ntaps * cin_p] in tap-major K order + tap offsets + bias."""

    def __init__(self, w: torch.Tensor, b: Optional[torch.Tensor], device, cin_pad: Optional[int] = None, cout_pad: Optional[int] = None,
                 taps: Optional[List[Tuple[int, int, int]]] = None):
        # w: [cout, cin, kt, kh, kw] (Conv3d) or [cout, cin, kh, kw] (Conv2d) or an explicit [cout, ntaps, cin] with `taps`
        if taps is None:
            if w.dim() == 4:
                w = w.unsqueeze(2)
            cout, cin, kt, kh, kw = w.shape
            taps = [(it - (kt - 1), ih - kh // 2, iw - kw // 2) for it in range(kt) for ih in range(kh) for iw in range(kw)]
            wm = w.permute(0, 2, 3, 4, 1).reshape(cout, kt * kh * kw, cin)
        else:
            wm = w
            cout, _, cin = wm.shape
        cin_p = cin_pad or _pad_to(cin, 32)
        cout_p = cout_pad or (16 if cout <= 16 else cout)
        full = torch.zeros(cout_p, len(taps), cin_p, dtype=torch.float32)
        full[:cout, :, :cin] = wm.float().cpu()
        self.weight = full.reshape(cout_p, len(taps) * cin_p).to(torch.bfloat16).to(device).contiguous()
        bias = torch.zeros(cout_p, dtype=torch.float32)
        if b is not None:
            bias[:cout] = b.float().cpu()
        self.bias = bias.to(torch.bfloat16).to(device)
        self.taps = taps
        self.taps_hist = [(dt + 2, dh, dw) for dt, dh, dw in taps]     # the same taps addressed into a view that starts two frames earlier
        self.temporal = any(dt != 0 for dt, _, _ in taps)
        self.cin, self.cout = cin_p, cout_p

    def __call__(self, x: torch.Tensor, out: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, clamp: bool = False):
        T, H, W, _ = x.shape
        if out is None:
            out = torch.empty((T, H, W, self.cout), dtype=torch.bfloat16, device=x.device)
        return lib.conv3d_cl(x, self.weight, self.bias, out, self.taps, residual=residual, clamp_out=clamp)

    def causal(self, buf: torch.Tensor, out: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, clamp: bool = False):
        """buf [T + 2, H, W, cin] = the two input frames that precede the chunk (zeros at the start of the sequence: the causal padding of
        CausalConv3d, vae.py:35-44; the cached frames of the reference's feat_cache afterwards, :203-217) followed by the chunk's T frames
        -> the convolution's output for those T frames."""
        Tp, H, W, _ = buf.shape
        if out is None:
            out = torch.empty((Tp - 2, H, W, self.cout), dtype=torch.bfloat16, device=buf.device)
        return lib.conv3d_cl_padded(buf, self.weight, self.bias, out, self.taps_hist, residual=residual, clamp_out=clamp)


class _UpsampleConv:
    """nn.Upsample(2x, nearest-exact) + Conv2d(3x3, padding 1) (vae.py:88-97) as four 2x2 phase convolutions."""

    def __init__(self, w: torch.Tensor, b: torch.Tensor, device):
        cout, cin = w.shape[:2]
        w = w.float().cpu()
        self.cout = cout
        self.phases = []
        for py in (0, 1):
            for px in (0, 1):
                # row taps: py=0 -> {y-1: w[0], y: w[1]+w[2]},  py=1 -> {y: w[0]+w[1], y+1: w[2]}   (likewise for columns)
                rows = [(-1, [0]), (0, [1, 2])] if py == 0 else [(0, [0, 1]), (1, [2])]
                cols = [(-1, [0]), (0, [1, 2])] if px == 0 else [(0, [0, 1]), (1, [2])]
                taps, mats = [], []
                for dh, khs in rows:
                    for dw, kws in cols:
                        taps.append((0, dh, dw))
                        mats.append(sum(w[:, :, kh, kw] for kh in khs for kw in kws))
                wm = torch.stack(mats, dim=1)                      # [cout, 4, cin]
                self.phases.append((py, px, _Conv(wm, b, device, taps=taps)))

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        T, H, W, _ = x.shape
        out = torch.empty((T, 2 * H, 2 * W, self.cout), dtype=torch.bfloat16, device=x.device)
        for py, px, conv in self.phases:
            conv(x, out=out[:, py::2, px::2])
        return out


class WanVAEDecoderB200:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", dim=96, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2,
                 temperal_upsample=(True, True, False), chunk_frames: Optional[int] = 3):
        W = state_dict
        self.device = torch.device(device)
        self.chunk_frames = chunk_frames           # latent frames decoded per pass (None: the whole sequence at once)
        self._hist: Dict[int, torch.Tensor] = {}   # id(conv) -> last two frames of that conv's input (chunked decode only)
        self._streaming = False
        self.z_dim = z_dim
        dev = self.device
        dims = [dim * u for u in [dim_mult[-1]] + list(dim_mult[::-1])]
        self.mean = torch.tensor(MEAN, dtype=torch.float32, device=dev)
        self.inv_std = 1.0 / torch.tensor(STD, dtype=torch.float32, device=dev)

        def g(name):
            return W[name].float().reshape(-1).to(dev).contiguous()

        def res(p, cin, cout):
            d = {"g0": g(p + ".residual.0.gamma"), "c1": _Conv(W[p + ".residual.2.weight"], W[p + ".residual.2.bias"], dev),
                 "g1": g(p + ".residual.3.gamma"), "c2": _Conv(W[p + ".residual.6.weight"], W[p + ".residual.6.bias"], dev), "sc": None}
            if cin != cout:
                d["sc"] = _Conv(W[p + ".shortcut.weight"], W[p + ".shortcut.bias"], dev)
            return d

        self.conv2 = _Conv(W["conv2.weight"], W["conv2.bias"], dev, cin_pad=32, cout_pad=64)     # 1x1x1, 16 -> 16 (zero-padded to 32 -> 64)
        self.conv1 = _Conv(W["decoder.conv1.weight"], W["decoder.conv1.bias"], dev, cin_pad=64)
        self.mid0 = res("decoder.middle.0", dims[0], dims[0])
        self.attn = {"g": g("decoder.middle.1.norm.gamma"),
                     "qkv": _Conv(W["decoder.middle.1.to_qkv.weight"], W["decoder.middle.1.to_qkv.bias"], dev),
                     "proj": _Conv(W["decoder.middle.1.proj.weight"], W["decoder.middle.1.proj.bias"], dev)}
        self.mid2 = res("decoder.middle.2", dims[0], dims[0])
        self.layers = []
        n = 0
        for i, (in_dim, out_dim) in enumerate(zip(dims[:-1], dims[1:])):
            if i in (1, 2, 3):
                in_dim = in_dim // 2
            for _ in range(num_res_blocks + 1):
                self.layers.append(("res", res(f"decoder.upsamples.{n}", in_dim, out_dim)))
                in_dim = out_dim
                n += 1
            if i != len(dim_mult) - 1:
                p = f"decoder.upsamples.{n}"
                up = {"conv": _UpsampleConv(W[p + ".resample.1.weight"], W[p + ".resample.1.bias"], dev), "time": None}
                if temperal_upsample[i]:
                    tw, tb = W[p + ".time_conv.weight"], W[p + ".time_conv.bias"]      # [2C, C, 3, 1, 1]
                    C = out_dim
                    up["time"] = [_Conv(tw[g_ * C:(g_ + 1) * C], tb[g_ * C:(g_ + 1) * C], dev) for g_ in (0, 1)]
                self.layers.append(("up", up))
                n += 1
        self.head_g = g("decoder.head.0.gamma")
        self.head = _Conv(W["decoder.head.2.weight"], W["decoder.head.2.bias"], dev, cout_pad=16)

    # ------------------------------------------------------------------ temporal convolutions with a two-frame history
    def _tconv(self, conv: _Conv, T: int, H: int, W: int, fill, **kw) -> torch.Tensor:
        """Run the causal convolution `conv` on a chunk of T frames whose input is produced by `fill(dst)` (dst = the [T, H, W, cin] slot the
        producer writes into - no concatenation copy).  Whole-sequence mode: plain zero-padded convolution.  Streaming mode: the two frames
        before the chunk come from the previous chunk (zeros for the first one) and the chunk's last two frames are kept for the next."""
        dev = self.device
        if not self._streaming:
            x = torch.empty((T, H, W, conv.cin), dtype=torch.bfloat16, device=dev)
            fill(x)
            return conv(x, **kw)
        buf = torch.empty((T + 2, H, W, conv.cin), dtype=torch.bfloat16, device=dev)
        prev = self._hist.get(id(conv))
        if prev is None:
            buf[:2].zero_()
        else:
            buf[:2].copy_(prev)
        fill(buf[2:])
        self._hist[id(conv)] = buf[-2:].clone()
        return conv.causal(buf, **kw)

    # ------------------------------------------------------------------ blocks
    def _res(self, d, x):
        T, H, W, _ = x.shape
        h = x if d["sc"] is None else d["sc"](x)
        a = self._tconv(d["c1"], T, H, W, lambda dst: lib.rms_silu_cl(x, d["g0"], out=dst))
        return self._tconv(d["c2"], T, H, W, lambda dst: lib.rms_silu_cl(a, d["g1"], out=dst), residual=h)

    def _attention(self, x):
        """AttentionBlock.forward (vae.py:245-262): per-frame single-head attention over the H*W positions, width C."""
        T, H, W, C = x.shape
        n = lib.rms_silu_cl(x, self.attn["g"], silu=False)
        qkv = self.attn["qkv"](n).view(T, H * W, 3, C)
        q, k, v = (qkv[:, :, i].unsqueeze(1) for i in range(3))          # [T, 1, HW, C]
        o = F.scaled_dot_product_attention(q, k, v).squeeze(1).reshape(T, H, W, C).contiguous()
        return self.attn["proj"](o, residual=x)

    def _upsample(self, up, x, first_chunk: bool = True):
        """Resample (vae.py:70-159).  upsample3d: the first latent frame of the SEQUENCE skips the temporal convolution and yields one
        frame ("Rep", :110-112); every later frame t yields two frames from time_conv over (f[t-2], f[t-1], f[t]) where the history
        starts at frame 1 (earlier frames read as zero).  `first_chunk` says whether x[0] is that first frame."""
        if up["time"] is not None:
            T, H, W, C = x.shape
            skip = 1 if first_chunk else 0                                # frames of this chunk that bypass the time_conv
            y = torch.empty((skip + 2 * (T - skip), H, W, C), dtype=torch.bfloat16, device=x.device)
            if skip:
                y[0].copy_(x[0])
            if T > skip:
                for g_, conv in enumerate(up["time"]):
                    self._tconv(conv, T - skip, H, W, lambda dst: dst.copy_(x[skip:]), out=y[skip + g_::2])
            elif self._streaming:
                for conv in up["time"]:                                   # a one-frame first chunk: the history of the time_conv stays empty
                    self._hist.pop(id(conv), None)
            x = y
        return up["conv"](x)

    # ------------------------------------------------------------------ reference surface
    @staticmethod
    def dist_slices(total: int, world_size: int, rank: int, padding: int = 1):
        """Strip bounds of WanVAE.decode_dist (vae.py:883-902): (latent slice taken by this rank, pixel crop of its decode).
        Rank 0 and the last rank take chunk + 2*padding columns from their end; middle ranks one halo column on each side."""
        chunk = total // world_size
        if rank == 0:
            return slice(0, chunk + 2 * padding), slice(0, chunk * 8)
        if rank == world_size - 1:
            return slice(total - (chunk + 2 * padding), total), slice(-chunk * 8, None)
        return slice(rank * chunk - padding, (rank + 1) * chunk + padding), slice(8 * padding, -8 * padding)

    @torch.no_grad()
    def decode_dist(self, zs: torch.Tensor, world_size: int, cur_rank: int, split_dim: int, group=None) -> torch.Tensor:
        """Width- or height-split parallel decode with a one-latent-pixel halo, crop, all_gather (vae.py:883-929).
        Like the reference this is an approximation of the unsplit decode near the seams (the receptive field exceeds the halo)."""
        import torch.distributed as dist

        lat, crop = self.dist_slices(zs.shape[split_dim], world_size, cur_rank)
        idx = [slice(None)] * 4
        idx[split_dim] = lat
        images = self._decode_local(zs[tuple(idx)].contiguous())
        cidx = [slice(None)] * 5
        cidx[split_dim + 1] = crop
        images = images[tuple(cidx)].contiguous()
        full = [torch.empty_like(images) for _ in range(world_size)]
        dist.all_gather(full, images, group=group)
        return torch.cat(full, dim=split_dim + 1)

    @torch.no_grad()
    def decode(self, zs: torch.Tensor, generator=None, config=None) -> torch.Tensor:
        """zs [16, T, H, W] fp32 -> [1, 3, 1+4(T-1), 8H, 8W] fp32 in [-1, 1] (WanVAE.decode, vae.py:931-957)."""
        if config is not None and config.get("parallel_vae", False):
            import torch.distributed as dist

            world, rank = dist.get_world_size(), dist.get_rank()
            if zs.shape[3] % world == 0:
                return self.decode_dist(zs, world, rank, 3)
            if zs.shape[2] % world == 0:
                return self.decode_dist(zs, world, rank, 2)
        return self._decode_local(zs)

    @torch.no_grad()
    def _decode_local(self, zs: torch.Tensor) -> torch.Tensor:
        zs = zs.to(self.device, torch.float32)
        Tl, Hl, Wl = zs.shape[1], zs.shape[2], zs.shape[3]
        chunk = self.chunk_frames
        if chunk is None or chunk >= Tl:
            self._streaming = False
            return lib.cl_to_video(self._decode_frames(zs, True)).unsqueeze(0)
        n_t = sum(1 for kind, layer in self.layers if kind == "up" and layer["time"] is not None)
        total = Tl
        for _ in range(n_t):
            total = 1 + 2 * (total - 1)
        video = torch.empty((3, total, Hl * 8, Wl * 8), dtype=torch.float32, device=self.device)
        self._streaming, self._hist = True, {}
        try:
            t_out = 0
            for t0 in range(0, Tl, chunk):
                x = self._decode_frames(zs[:, t0:t0 + chunk].contiguous(), t0 == 0)
                lib.cl_to_video(x, out=video[:, t_out:t_out + x.shape[0]])
                t_out += x.shape[0]
                del x
            assert t_out == total, (t_out, total)
        finally:
            self._streaming, self._hist = False, {}
        return video.unsqueeze(0)

    def _decode_frames(self, zs: torch.Tensor, first_chunk: bool) -> torch.Tensor:
        """Latent frames [16, T, H, W] fp32 -> channels-last bf16 video frames [T', 8H, 8W, 16] (3 valid channels, clamped to [-1, 1])."""
        x = lib.latent_to_cl(zs, self.mean, self.inv_std, cp=32)
        T, H, W, _ = x.shape
        x = self._tconv(self.conv1, T, H, W, lambda dst: self.conv2(x, out=dst))
        x = self._res(self.mid0, x)
        x = self._attention(x)
        x = self._res(self.mid2, x)
        for kind, layer in self.layers:
            x = self._res(layer, x) if kind == "res" else self._upsample(layer, x, first_chunk)
        T, H, W, _ = x.shape
        return self._tconv(self.head, T, H, W, lambda dst: lib.rms_silu_cl(x, self.head_g, out=dst), clamp=True)


# ============================================================================================================================
# Encoder (image-to-video conditioning: WanVAE.encode, vae.py:867-880 -> WanVAE_.encode :684-711 -> Encoder3d :264-376)
# ============================================================================================================================
class _DownConv:
    """nn.ZeroPad2d((0, 1, 0, 1)) + nn.Conv2d(dim, dim, 3, stride 2) (vae.py:99-104) on the stride-1 implicit-GEMM kernel: the input is
    read through its four (row, column) parity views x[:, ph::2, pw::2] (strides doubled, no copy); in view coordinates the nine taps
    become 4 + 2 + 2 + 1 unit-stride taps, issued as four launches that accumulate through the residual epilogue.  Reads past the end
    of a view return zero, which is the bottom / right zero padding."""

    def __init__(self, w: torch.Tensor, b: torch.Tensor, device):
        w = w.float().cpu()                                   # [cout, cin, 3, 3]
        self.cout = w.shape[0]
        self.parts = []
        first = True
        for ph in (0, 1):
            for pw in (0, 1):
                taps, mats = [], []
                for dh in range(ph, 3, 2):
                    for dw in range(pw, 3, 2):
                        taps.append((0, dh // 2, dw // 2))
                        mats.append(w[:, :, dh, dw])
                self.parts.append((ph, pw, _Conv(torch.stack(mats, dim=1), b if first else None, device, taps=taps)))
                first = False

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        T, H, W, _ = x.shape
        Ho, Wo = (H + 1 - 3) // 2 + 1, (W + 1 - 3) // 2 + 1
        out = torch.empty((T, Ho, Wo, self.cout), dtype=torch.bfloat16, device=x.device)
        for i, (ph, pw, conv) in enumerate(self.parts):
            lib.conv3d_cl_padded(x[:, ph::2, pw::2], conv.weight, conv.bias, out, conv.taps, residual=out if i else None)
        return out


class _TimeDownConv:
    """CausalConv3d(dim, dim, (3, 1, 1), stride (2, 1, 1), padding 0) as the reference applies it across its 1, 4, 4, ... frame chunks
    (Resample.forward downsample3d, vae.py:144-158): y[0] = x[0]; y[k] = W0 x[2k-2] + W1 x[2k-1] + W2 x[2k] for k >= 1.  Two launches on
    the even / odd frame views: y[1 + k'] = W0 x_even[k'] + W2 x_even[k' + 1]  (+)  W1 x_odd[k']."""

    def __init__(self, w: torch.Tensor, b: torch.Tensor, device):
        w = w.float().cpu()                                   # [cout, cin, 3, 1, 1]
        self.even = _Conv(torch.stack([w[:, :, 0, 0, 0], w[:, :, 2, 0, 0]], dim=1), b, device, taps=[(0, 0, 0), (1, 0, 0)])
        self.odd = _Conv(w[:, :, 1, 0, 0].unsqueeze(1), None, device, taps=[(0, 0, 0)])

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        T, H, W, C = x.shape
        if T % 2 != 1:
            raise lib.B200Error(f"WanVAE encoder: temporal downsampling needs an odd frame count (1 + 2k), got {T}")
        To = 1 + (T - 1) // 2
        y = torch.empty((To, H, W, self.even.cout), dtype=torch.bfloat16, device=x.device)
        y[0].copy_(x[0])
        if To > 1:
            lib.conv3d_cl_padded(x[0::2], self.even.weight, self.even.bias, y[1:], self.even.taps)
            lib.conv3d_cl_padded(x[1::2], self.odd.weight, self.odd.bias, y[1:], self.odd.taps, residual=y[1:])
        return y


class WanVAEEncoderB200:
    """`encode(video [3, T, H, W] in [-1, 1], T = 1 + 4k) -> normalised latent mean [16, 1 + k, H/8, W/8] fp32`
    (WanVAE.encode, vae.py:867-880; the reference feeds 1, 4, 4, ... frames with per-convolution caches, which is the whole-sequence
    causal convolution computed here).  Same kernels and layer helpers as the decoder; activations bf16 (the reference: fp32)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", dim=96, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2,
                 temperal_downsample=(False, True, True)):
        W = state_dict
        self.device = dev = torch.device(device)
        self.z_dim = z_dim
        self.mean = torch.tensor(MEAN, dtype=torch.float32, device=dev)
        self.inv_std = 1.0 / torch.tensor(STD, dtype=torch.float32, device=dev)
        self._zero3, self._one3 = torch.zeros(3, device=dev), torch.ones(3, device=dev)

        def g(name):
            return W[name].float().reshape(-1).to(dev).contiguous()

        def res(p, cin, cout):
            d = {"g0": g(p + ".residual.0.gamma"), "c1": _Conv(W[p + ".residual.2.weight"], W[p + ".residual.2.bias"], dev),
                 "g1": g(p + ".residual.3.gamma"), "c2": _Conv(W[p + ".residual.6.weight"], W[p + ".residual.6.bias"], dev), "sc": None}
            if cin != cout:
                d["sc"] = _Conv(W[p + ".shortcut.weight"], W[p + ".shortcut.bias"], dev)
            return d

        dims = [dim * u for u in [1] + list(dim_mult)]
        self.conv_in = _Conv(W["encoder.conv1.weight"], W["encoder.conv1.bias"], dev, cin_pad=32)       # 3 -> 96 (input zero-padded to 32 channels)
        self.layers = []
        n, out_dim = 0, dims[0]
        for i, (in_dim, out_dim) in enumerate(zip(dims[:-1], dims[1:])):
            for _ in range(num_res_blocks):
                self.layers.append(("res", res(f"encoder.downsamples.{n}", in_dim, out_dim)))
                in_dim = out_dim
                n += 1
            if i != len(dim_mult) - 1:
                p = f"encoder.downsamples.{n}"
                down = {"conv": _DownConv(W[p + ".resample.1.weight"], W[p + ".resample.1.bias"], dev), "time": None}
                if temperal_downsample[i]:
                    down["time"] = _TimeDownConv(W[p + ".time_conv.weight"], W[p + ".time_conv.bias"], dev)
                self.layers.append(("down", down))
                n += 1
        self.mid0 = res("encoder.middle.0", out_dim, out_dim)
        self.attn = {"g": g("encoder.middle.1.norm.gamma"),
                     "qkv": _Conv(W["encoder.middle.1.to_qkv.weight"], W["encoder.middle.1.to_qkv.bias"], dev),
                     "proj": _Conv(W["encoder.middle.1.proj.weight"], W["encoder.middle.1.proj.bias"], dev)}
        self.mid2 = res("encoder.middle.2", out_dim, out_dim)
        self.head_g = g("encoder.head.0.gamma")
        self.head = _Conv(W["encoder.head.2.weight"], W["encoder.head.2.bias"], dev, cout_pad=64)        # 384 -> 2 z_dim = 32 (padded to 64)
        self.conv1 = _Conv(W["conv1.weight"], W["conv1.bias"], dev, cin_pad=64, cout_pad=64)            # 1x1x1 in front of the mu / log_var split

    @staticmethod
    def _res(d, x):
        h = x if d["sc"] is None else d["sc"](x)
        a = d["c1"](lib.rms_silu_cl(x, d["g0"]))
        return d["c2"](lib.rms_silu_cl(a, d["g1"]), residual=h)

    _attention = WanVAEDecoderB200._attention

    @torch.no_grad()
    def encode(self, video: torch.Tensor) -> torch.Tensor:
        video = video.to(self.device, torch.float32)
        if video.dim() == 5:
            video = video[0]
        x = lib.latent_to_cl(video.contiguous(), self._zero3, self._one3, cp=32)      # [3, T, H, W] fp32 -> channels-last bf16, zero-padded channels
        x = self.conv_in(x)
        for kind, layer in self.layers:
            if kind == "res":
                x = self._res(layer, x)
            else:
                x = layer["conv"](x)
                if layer["time"] is not None:
                    x = layer["time"](x)
        x = self._res(self.mid0, x)
        x = self._attention(x)
        x = self._res(self.mid2, x)
        x = self.head(lib.rms_silu_cl(x, self.head_g))
        x = self.conv1(x)                                                               # [T', H/8, W/8, 64]: channels [0, 16) = mu
        mu = x[..., : self.z_dim].permute(3, 0, 1, 2).float()
        return (mu - self.mean.view(-1, 1, 1, 1)) * self.inv_std.view(-1, 1, 1, 1)
