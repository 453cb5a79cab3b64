"""ORACLE (test infrastructure, NOT product code).

A plain-torch restatement of LightX2V's Wan DiT hot path in its `DTYPE=BF16` mode, function by function, each citing
the reference lines it follows (paths relative to the LightX2V tree at 0591c35e).  It runs on CPU (attention through
torch SDPA, like the reference's `torch_sdpa` op) or on a GPU (`device="cuda"`, attention through flash_attn when
`attn="flash_attn2"`, like the reference's `flash_attn2` op) and is the checker for every CUDA kernel in
`lightx2v_b200/csrc`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this file.

Pinning: `oracle/gen_golden.py` imports the REAL reference classes (under the two shims of SURVEY.md §8c) in the build
container, runs them on seeded synthetic weights and stores inputs + outputs under `tests/golden/`;
`tests/test_oracle_golden.py` checks this restatement against those fixtures bit for bit.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------------------------
# operator surface (lightx2v/common/ops/*)
# ---------------------------------------------------------------------------------------------------------------
def mm_apply(x: torch.Tensor, weight_nk: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """MMWeight.apply: torch.addmm(bias, x, W.t())  — common/ops/mm/mm_weight.py:81-88 (weight kept as [N,K].t() view, :76)."""
    if bias is None:
        return torch.mm(x, weight_nk.t())
    return torch.addmm(bias, x, weight_nk.t())


# ---- w8a8-fp8 linear (BASELINE config 3) ------------------------------------------------------------------------
def fp8_weight_quant(w: torch.Tensor):
    """FloatQuantizer("e4m3", True, "per_channel").real_quant_tensor — utils/quant_utils.py:41-53 (absmax.clamp(1e-5)/qmax),
    :149-161 (clip to +-448, round to nearest e4m3); used by load_fp8_perchannel_sym, mm_weight.py:167-183."""
    wf = w.to(torch.float32)
    scale = wf.abs().amax(dim=-1, keepdim=True).clamp(min=1e-5) / 448.0
    q = torch.clip(wf / scale, -448.0, 448.0).to(torch.float8_e4m3fn)
    return q, scale


def fp8_act_quant(x: torch.Tensor):
    """vLLM ops.scaled_fp8_quant(x, None, scale_ub=None, use_per_token_if_dynamic=True) — mm_weight.py:236-238.
    vLLM is a third-party dependency absent from /root/reference (docs pin 0.9.2; the image has 0.22): published algorithm of
    its dynamic_per_token_scaled_fp8_quant kernel: scale = max(absmax/448, 1/(448*512)); q = sat_e4m3_rn(x / scale)."""
    xf = x.to(torch.float32)
    # true fp32 division (a tensor divisor: torch turns `tensor / python_scalar` into a multiply by the reciprocal on CUDA)
    scale = torch.div(xf.abs().amax(dim=-1, keepdim=True), torch.full((), 448.0, device=xf.device)).clamp(min=1.0 / (448.0 * 512.0))
    q = torch.clip(xf / scale, -448.0, 448.0).to(torch.float8_e4m3fn)
    return q, scale


def mm_fp8_apply(x: torch.Tensor, wq: torch.Tensor, wscale: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """MMWeightWfp8channelAfp8channeldynamicVllm.apply — mm_weight.py:304-319: act quant, then cutlass_scaled_mm:
    out = a_scale * (b_scale * (xq @ wq^T)) + bias in fp32, rounded once to bf16."""
    xq, sx = fp8_act_quant(x)
    acc = xq.to(torch.float32) @ wq.to(torch.float32).t()
    y = sx * (wscale.reshape(1, -1) * acc)
    if bias is not None:
        y = y + bias.to(torch.float32)
    return y.to(torch.bfloat16)


def mm_named(W: Dict[str, torch.Tensor], name: str, x: torch.Tensor) -> torch.Tensor:
    """Linear `name` of a checkpoint-named dict: bf16 MMWeight, or the w8a8-fp8 class when `<name>.weight_scale` is present
    (the reference picks the class from mm_config.mm_type, transformer_weights.py:20-23; the quantised checkpoint carries the scales)."""
    if name + ".weight_global_scale" in W:
        return mm_nvfp4_apply(x, W[name + ".weight"], W.get(name + ".bias"))
    if name + ".weight_scale" in W:
        return mm_fp8_apply(x, W[name + ".weight"], W[name + ".weight_scale"], W.get(name + ".bias"))
    return mm_apply(x, W[name + ".weight"], W.get(name + ".bias"))


def nvfp4_fake_quant(x: torch.Tensor) -> torch.Tensor:
    """x -> dequantise(quantise(x)) in fp32 with the dynamic per-tensor global scale 448*6/max|x| - the reference's NVFP4 golden
    model (lightx2v_kernel/test/nvfp4_nvfp4/fake_quant.py:34-51 + test_bench1.py:57-72), see oracle/nvfp4_oracle.py."""
    from . import nvfp4_oracle as NV

    gs = (NV.E4M3_MAX * NV.E2M1_MAX / x.float().abs().max()).to(torch.float32)
    vals, scale = NV.quant(x, gs)
    m, k = x.shape
    return (vals.reshape(m, k // 16, 16) * (scale / gs).unsqueeze(-1)).reshape(m, k)


def mm_nvfp4_apply(x: torch.Tensor, w_dq: torch.Tensor, bias) -> torch.Tensor:
    """w4a4 linear as the reference's GEMM test defines it (test_bench1.py:75-103,134): dequantised operands, fp32 matmul, + bias, -> bf16.
    `w_dq` is the already fake-quantised fp32 weight (quantize_checkpoint_nvfp4)."""
    y = nvfp4_fake_quant(x) @ w_dq.t()
    if bias is not None:
        y = y + bias.float()
    return y.to(torch.bfloat16)


def quantize_checkpoint_nvfp4(W: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Oracle-side w4a4 conversion of the block linears: `<name>.weight` -> fake-quantised fp32 values + a `<name>.weight_global_scale`
    marker (the product-side converter, lightx2v_b200.host.ops.quantize_checkpoint_nvfp4, emits the packed format instead)."""
    out = dict(W)
    for k, v in W.items():
        if k.endswith(".weight") and v.dim() == 2 and ("attn." in k or "ffn." in k) and "norm" not in k:
            out[k] = nvfp4_fake_quant(v)
            out[k[: -len(".weight")] + ".weight_global_scale"] = (448.0 * 6.0 / v.float().abs().max()).reshape(1)
    return out


def quantize_checkpoint_fp8(W: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Offline w8a8 conversion of the block linears (tools/convert/converter.py:294-407): `<name>.weight` -> e4m3 + `<name>.weight_scale`."""
    out = dict(W)
    for k, v in W.items():
        if k.endswith(".weight") and v.dim() == 2 and ("attn." in k or "ffn." in k) and "norm" not in k:
            q, s = fp8_weight_quant(v)
            out[k] = q
            out[k[: -len(".weight")] + ".weight_scale"] = s
    return out


def psnr(got: torch.Tensor, ref: torch.Tensor) -> float:
    """PSNR in dB with the reference's peak as signal (north_star: fp8 / nvfp4 paths report PSNR vs the bf16 reference)."""
    got, ref = got.float(), ref.float()
    mse = (got - ref).pow(2).mean().clamp(min=1e-30)
    return float(10.0 * torch.log10(ref.abs().max().pow(2) / mse))


def rms_apply(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """RMSWeightSgl.apply, bf16 fallback taken when sgl_kernel is absent — common/ops/norm/rms_norm_weight.py:109-113."""
    x = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    return x * weight


def ln_apply(x: torch.Tensor, weight=None, bias=None, eps: float = 1e-6) -> torch.Tensor:
    """LNWeight.apply in BF16 mode — common/ops/norm/layer_norm_weight.py:110."""
    return F.layer_norm(x, (x.shape[-1],), weight, bias, eps)


def attn_apply(q, k, v, impl: str = "torch_sdpa") -> torch.Tensor:
    """ATTN op `.apply` for one segment, non-causal, scale d^-0.5 -> [Sq, H*d].
    torch_sdpa: common/ops/attn/attn_weight.py:209-239;  flash_attn2: :71-97."""
    sq = q.shape[0]
    if impl == "flash_attn2":
        from flash_attn import flash_attn_varlen_func

        cu_q = torch.tensor([0, sq], dtype=torch.int32, device=q.device)
        cu_k = torch.tensor([0, k.shape[0]], dtype=torch.int32, device=q.device)
        return flash_attn_varlen_func(q, k, v, cu_q, cu_k, sq, k.shape[0]).reshape(sq, -1)
    qq, kk, vv = (t.unsqueeze(0).transpose(1, 2) for t in (q, k, v))
    x = F.scaled_dot_product_attention(qq, kk, vv, attn_mask=None, dropout_p=0, is_causal=False)
    return x.transpose(1, 2).reshape(1, sq, -1).squeeze(0)


# ---------------------------------------------------------------------------------------------------------------
# RoPE (models/networks/wan/infer/utils.py)
# ---------------------------------------------------------------------------------------------------------------
def rope_params(max_seq_len: int, dim: int, theta: float = 10000.0) -> torch.Tensor:
    """utils.py:151-158 — complex128 table [max_seq_len, dim/2]."""
    freqs = torch.outer(torch.arange(max_seq_len), 1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim)))
    return torch.polar(torch.ones_like(freqs), freqs)


def wan_freqs_table(head_dim: int = 128) -> torch.Tensor:
    """WanModel / WanPreInfer build: cat of three rope_params tables (pre_infer.py:16-25) -> [1024, head_dim/2] complex128."""
    d = head_dim
    return torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)), rope_params(1024, 2 * (d // 6))], dim=1)


def compute_freqs(c: int, grid_sizes, freqs: torch.Tensor) -> torch.Tensor:
    """utils.py:7-20 — [f*h*w, 1, c] complex table for the (t, h, w) grid."""
    fs = freqs.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    f, h, w = [int(v) for v in grid_sizes]
    return torch.cat(
        [
            fs[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1),
            fs[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
            fs[2][:w].view(1, 1, w, -1).expand(f, h, w, -1),
        ],
        dim=-1,
    ).reshape(f * h * w, 1, -1)


def compute_freqs_dist(s: int, c: int, grid_sizes, freqs: torch.Tensor, world_size: int, rank: int) -> torch.Tensor:
    """utils.py:86-104 — pad with ones (identity rotation) to s*world_size rows and slice this rank's rows."""
    fi = compute_freqs(c, grid_sizes, freqs)
    pad = s * world_size - fi.shape[0]
    if pad > 0:
        fi = torch.cat([fi, torch.ones(pad, 1, fi.shape[2], dtype=fi.dtype, device=fi.device)], dim=0)
    return fi[rank * s : (rank + 1) * s]


def apply_rotary_emb(x: torch.Tensor, freqs_i: torch.Tensor) -> torch.Tensor:
    """utils.py:107-115 — complex128 multiply of adjacent pairs, single rounding to bf16."""
    n = x.size(1)
    seq_len = freqs_i.size(0)
    x_i = torch.view_as_complex(x[:seq_len].to(torch.float64).reshape(seq_len, n, -1, 2))
    x_i = torch.view_as_real(x_i * freqs_i).flatten(2)
    x_i = torch.cat([x_i, x[seq_len:]])
    return x_i.to(torch.float64 if x.dtype == torch.float64 else torch.bfloat16)   # fp64 only in infer_blocks_exact (not a reference mode)


def cos_sin_table(freqs_i: torch.Tensor) -> torch.Tensor:
    """Host-side product helper input format: [S, 64, 2] fp32 (cos, sin) from the complex128 table (rounded once)."""
    f = freqs_i.reshape(freqs_i.shape[0], -1)
    return torch.stack([f.real, f.imag], dim=-1).to(torch.float32).contiguous()


# ---------------------------------------------------------------------------------------------------------------
# one DiT block  (models/networks/wan/infer/transformer_infer.py:289-508), weights = checkpoint-named dict
# ---------------------------------------------------------------------------------------------------------------
def infer_modulation(W: Dict[str, torch.Tensor], pre: str, embed0: torch.Tensor):
    """transformer_infer.py:308-319 (embed0.dim()==2 branch: [6, D]; modulation [1, 6, D])."""
    return (W[pre + "modulation"] + embed0).chunk(6, dim=1)


def infer_self_attn(W, pre, x, freqs_i, shift_msa, scale_msa, num_heads, attn="torch_sdpa", parallel_attention=None):
    """transformer_infer.py:321-396."""
    norm1_weight = 1 + scale_msa.squeeze(0)
    norm1_bias = shift_msa.squeeze(0)
    norm1_out = ln_apply(x)                                              # :329 (norm1: LN without affine)
    norm1_out.mul_(norm1_weight).add_(norm1_bias)                        # :334
    s, n = norm1_out.shape[0], num_heads
    d = norm1_out.shape[1] // n
    sa = pre + "self_attn."
    q = rms_apply(mm_named(W, sa + "q", norm1_out), W[sa + "norm_q.weight"]).view(s, n, d)   # :341
    k = rms_apply(mm_named(W, sa + "k", norm1_out), W[sa + "norm_k.weight"]).view(s, n, d)   # :342
    v = mm_named(W, sa + "v", norm1_out).view(s, n, d)                                       # :343
    q = apply_rotary_emb(q, freqs_i)                                     # :358
    k = apply_rotary_emb(k, freqs_i)                                     # :359
    if parallel_attention is None:
        attn_out = attn_apply(q, k, v, attn)                             # :369-379
    else:
        attn_out = parallel_attention(q, k, v)                           # :381-388
    return mm_named(W, sa + "o", attn_out)      # :390


def infer_cross_attn(W, pre, x, context, y_out, gate_msa, num_heads, task="t2v", attn="torch_sdpa"):
    """transformer_infer.py:398-465."""
    x.add_(y_out * gate_msa.squeeze(0))                                  # :402
    ca = pre + "cross_attn."
    norm3_out = ln_apply(x, W[pre + "norm3.weight"], W[pre + "norm3.bias"])   # :404
    if task == "i2v":
        context_img, context = context[:257], context[257:]             # :405-407
    n = num_heads
    d = x.shape[1] // n
    q = rms_apply(mm_named(W, ca + "q", norm3_out), W[ca + "norm_q.weight"]).view(-1, n, d)  # :418
    k = rms_apply(mm_named(W, ca + "k", context), W[ca + "norm_k.weight"]).view(-1, n, d)    # :419
    v = mm_named(W, ca + "v", context).view(-1, n, d)                                        # :420
    attn_out = attn_apply(q, k, v, attn)                                 # :425-434
    if task == "i2v":
        k_img = rms_apply(mm_named(W, ca + "k_img", context_img), W[ca + "norm_k_img.weight"]).view(-1, n, d)
        v_img = mm_named(W, ca + "v_img", context_img).view(-1, n, d)
        attn_out = attn_out.clone() if not attn_out.is_contiguous() else attn_out
        attn_out.add_(attn_apply(q, k_img, v_img, attn))                 # :436-454
    attn_out = mm_named(W, ca + "o", attn_out)  # :460
    return x, attn_out


def infer_ffn(W, pre, x, attn_out, c_shift_msa, c_scale_msa):
    """transformer_infer.py:467-497."""
    x.add_(attn_out)                                                     # :468
    norm2_weight = 1 + c_scale_msa.squeeze(0)
    norm2_bias = c_shift_msa.squeeze(0)
    norm2_out = ln_apply(x)                                              # :481
    norm2_out.mul_(norm2_weight).add_(norm2_bias)                        # :484
    y = mm_named(W, pre + "ffn.0", norm2_out)   # :488
    y = F.gelu(y, approximate="tanh")                                    # :492
    return mm_named(W, pre + "ffn.2", y)   # :495


def post_process(x, y, c_gate_msa):
    """transformer_infer.py:499-508."""
    x.add_(y * c_gate_msa.squeeze(0))                                    # :503
    return x


def infer_block(W, block_idx, x, embed0, freqs_i, context, num_heads, task="t2v", attn="torch_sdpa", parallel_attention=None):
    """WanTransformerInfer.infer_block — transformer_infer.py:289-306.  Mutates x in place like the reference."""
    pre = f"blocks.{block_idx}."
    shift_msa, scale_msa, gate_msa, c_shift_msa, c_scale_msa, c_gate_msa = infer_modulation(W, pre, embed0)
    y_out = infer_self_attn(W, pre, x, freqs_i, shift_msa, scale_msa, num_heads, attn, parallel_attention)
    x, attn_out = infer_cross_attn(W, pre, x, context, y_out, gate_msa, num_heads, task, attn)
    y = infer_ffn(W, pre, x, attn_out, c_shift_msa, c_scale_msa)
    return post_process(x, y, c_gate_msa)


def infer_blocks(W, num_layers, x, embed0, grid_sizes, freqs, context, num_heads, task="t2v", attn="torch_sdpa"):
    """_infer_without_offload — transformer_infer.py:269-287 (freqs_i recomputed per block in the reference, :349)."""
    d = x.shape[1] // num_heads
    for i in range(num_layers):
        freqs_i = compute_freqs(d // 2, grid_sizes, freqs)
        x = infer_block(W, i, x, embed0, freqs_i, context, num_heads, task, attn)
    return x


def infer_blocks_exact(W, num_layers, x, embed0, grid_sizes, freqs, context, num_heads, task="t2v"):
    """The same block stack evaluated in float64 from the same bf16 weights and inputs, with NO intermediate rounding: the "truth" both the
    reference's bf16 path and the CUDA path approximate.  Not a reference mode - used by the tests that show the CUDA path is as close to
    the exact result as the reference's own bf16 evaluation is (the honest way to read an element-wise tolerance between two bf16 pipelines)."""
    Wd = {k: v.to(torch.float64) for k, v in W.items()}
    return infer_blocks(Wd, num_layers, x.to(torch.float64), embed0.to(torch.float64), grid_sizes, freqs.to(x.device), context.to(torch.float64), num_heads,
                        task, "torch_sdpa")


# ---------------------------------------------------------------------------------------------------------------
# pre-infer / post-infer (A13) and the CFG combine (A18)
# ---------------------------------------------------------------------------------------------------------------
def sinusoidal_embedding_1d(dim: int, position: torch.Tensor) -> torch.Tensor:
    """utils.py:161-172 (BF16 mode)."""
    half = dim // 2
    position = position.type(torch.float64)
    sinusoid = torch.outer(position, torch.pow(10000, -torch.arange(half).to(position).div(half)))
    return torch.cat([torch.cos(sinusoid), torch.sin(sinusoid)], dim=1).to(torch.bfloat16)


def pre_infer(W, latents, t, context, dim: int, freq_dim: int = 256, text_len: int = 512, clip_fea=None, vae_encode_out=None):
    """WanPreInfer.infer - pre_infer.py:29-120 (no diffusion forcing, no dynamic cfg, seq_len == token count).
    latents [C, F, H, W] bf16; t [1]; context [L <= 512, 4096].  Returns embed [1, D], grid (F, H/2, W/2), x [S, D], embed0 [6, D], context'."""
    x = latents
    if vae_encode_out is not None:
        x = torch.cat([x, vae_encode_out.to(x.dtype)], dim=0)                                         # :52-53
    u = F.conv3d(x.unsqueeze(0), W["patch_embedding.weight"], W["patch_embedding.bias"], stride=(1, 2, 2))   # :56
    grid = tuple(int(v) for v in u.shape[2:])
    xs = u.flatten(2).transpose(1, 2)[0]
    embed = sinusoidal_embedding_1d(freq_dim, t.flatten())
    embed = mm_named(W, "time_embedding.0", embed)
    embed = mm_named(W, "time_embedding.2", F.silu(embed))
    embed0 = mm_named(W, "time_projection.1", F.silu(embed)).unflatten(1, (6, dim)).squeeze(0)       # :75-78
    ctx = torch.cat([context, context.new_zeros(text_len - context.size(0), context.size(1))])       # :90
    ctx = mm_named(W, "text_embedding.2", F.gelu(mm_named(W, "text_embedding.0", ctx), approximate="tanh"))
    if clip_fea is not None:                                                                         # :100-111
        c = ln_apply(clip_fea, W["img_emb.proj.0.weight"], W["img_emb.proj.0.bias"], 1e-6)
        c = F.gelu(mm_named(W, "img_emb.proj.1", c), approximate="none")
        c = ln_apply(mm_named(W, "img_emb.proj.3", c), W["img_emb.proj.4.weight"], W["img_emb.proj.4.bias"], 1e-6)
        ctx = torch.cat([c, ctx], dim=0)
    return embed, grid, xs, embed0, ctx


def post_infer(W, x, e, grid, out_dim: int = 16):
    """WanPostInfer.infer - post_infer.py:15-50 (BF16 mode)."""
    e0, e1 = (W["head.modulation"] + e.unsqueeze(1)).chunk(2, dim=1)
    x = ln_apply(x)
    x = x.mul_(1 + e1.squeeze(0)).add_(e0.squeeze(0))
    x = mm_named(W, "head.head", x)
    f, h, w = grid
    u = x[: f * h * w].view(f, h, w, 1, 2, 2, out_dim)
    u = torch.einsum("fhwpqrc->cfphqwr", u)
    return u.reshape(out_dim, f, h * 2, w * 2).float()


def cfg_combine(cond, uncond, scale: float):
    """WanModel.infer - wan/model.py:216-218."""
    return uncond + scale * (cond - uncond)


def synth_prepost_weights(dim: int, in_dim: int = 16, task: str = "t2v", seed: int = 13, device="cpu") -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    W: Dict[str, torch.Tensor] = {}

    def lin(name, n, k, scale=None):
        W[name + ".weight"] = (torch.randn(n, k, generator=g) * (scale or k ** -0.5)).to(torch.bfloat16).to(device)
        W[name + ".bias"] = (torch.randn(n, generator=g) * 0.02).to(torch.bfloat16).to(device)

    W["patch_embedding.weight"] = (torch.randn(dim, in_dim, 1, 2, 2, generator=g) * (in_dim * 4) ** -0.5).to(torch.bfloat16).to(device)
    W["patch_embedding.bias"] = (torch.randn(dim, generator=g) * 0.02).to(torch.bfloat16).to(device)
    lin("text_embedding.0", dim, 4096)
    lin("text_embedding.2", dim, dim)
    lin("time_embedding.0", dim, 256)
    lin("time_embedding.2", dim, dim)
    lin("time_projection.1", 6 * dim, dim)
    lin("head.head", 64, dim)
    W["head.modulation"] = (torch.randn(1, 2, dim, generator=g) * 0.1).to(torch.bfloat16).to(device)
    if task == "i2v":
        for n, c in (("img_emb.proj.0", 1280), ("img_emb.proj.4", dim)):
            W[n + ".weight"] = (1 + 0.05 * torch.randn(c, generator=g)).to(torch.bfloat16).to(device)
            W[n + ".bias"] = (0.02 * torch.randn(c, generator=g)).to(torch.bfloat16).to(device)
        lin("img_emb.proj.1", 1280, 1280)
        lin("img_emb.proj.3", dim, 1280)
    return W


# ---------------------------------------------------------------------------------------------------------------
# CausVid: autoregressive block variant with a self-attention KV cache (wan/infer/causvid/transformer_infer.py)
# ---------------------------------------------------------------------------------------------------------------
def compute_freqs_causvid(c: int, grid_sizes, freqs: torch.Tensor, start_frame: int = 0) -> torch.Tensor:
    """utils.py:62-75 - like compute_freqs but the temporal rows start at `start_frame`."""
    fs = freqs.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    f, h, w = [int(v) for v in grid_sizes]
    return torch.cat(
        [
            fs[0][start_frame:start_frame + f].view(f, 1, 1, -1).expand(f, h, w, -1),
            fs[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
            fs[2][:w].view(1, 1, w, -1).expand(f, h, w, -1),
        ],
        dim=-1,
    ).reshape(f * h * w, 1, -1)


def infer_block_causvid(W, block_idx, x, embed0, grid_sizes, freqs, context, num_heads, kv_cache, kv_start, kv_end, attn="torch_sdpa"):
    """WanTransformerInferCausVid.infer_block (t2v) - causvid/transformer_infer.py:205-220: self-attention :95-137 (q/k/v of the chunk,
    RoPE at the chunk's frame offset, K/V appended to the cache at [kv_start, kv_end), attention of the chunk's queries against
    cache[:kv_end]), cross-attention :139-190, FFN :192-199.  kv_cache: dict(k=[N, H, d], v=[N, H, d]) of this block, updated in place."""
    pre = f"blocks.{block_idx}."
    e = (W[pre + "modulation"] + embed0).chunk(6, dim=1)                                  # :211-212 (embed0.dim() == 2)
    norm1_out = ln_apply(x)
    norm1_out = (norm1_out * (1 + e[1]) + e[0]).squeeze(0)                                # :96-97
    s, n = norm1_out.shape[0], num_heads
    d = norm1_out.shape[1] // n
    sa = pre + "self_attn."
    q = rms_apply(mm_named(W, sa + "q", norm1_out), W[sa + "norm_q.weight"]).view(s, n, d)
    k = rms_apply(mm_named(W, sa + "k", norm1_out), W[sa + "norm_k.weight"]).view(s, n, d)
    v = mm_named(W, sa + "v", norm1_out).view(s, n, d)
    f, h, w = [int(t) for t in grid_sizes]
    freqs_i = compute_freqs_causvid(d // 2, grid_sizes, freqs, start_frame=kv_start // (h * w))   # :104
    q, k = apply_rotary_emb(q, freqs_i), apply_rotary_emb(k, freqs_i)
    kv_cache["k"][kv_start:kv_end] = k                                                    # :112-113
    kv_cache["v"][kv_start:kv_end] = v
    attn_out = attn_apply(q, kv_cache["k"][:kv_end], kv_cache["v"][:kv_end], attn)        # :118-127
    x = x + mm_named(W, sa + "o", attn_out) * e[2].squeeze(0)                             # :133-135
    # cross-attention (:139-190): identical maths to the base class, K/V of the prompt cached after the first chunk
    ca = pre + "cross_attn."
    norm3_out = ln_apply(x, W[pre + "norm3.weight"], W[pre + "norm3.bias"])
    cq = rms_apply(mm_named(W, ca + "q", norm3_out), W[ca + "norm_q.weight"]).view(-1, n, d)
    ck = rms_apply(mm_named(W, ca + "k", context), W[ca + "norm_k.weight"]).view(-1, n, d)
    cv = mm_named(W, ca + "v", context).view(-1, n, d)
    x = x + mm_named(W, ca + "o", attn_apply(cq, ck, cv, attn))
    # FFN (:192-199)
    norm2_out = ln_apply(x)
    y = mm_named(W, pre + "ffn.0", norm2_out * (1 + e[4].squeeze(0)) + e[3].squeeze(0))
    y = mm_named(W, pre + "ffn.2", F.gelu(y, approximate="tanh"))
    return x + y * e[5].squeeze(0)


def infer_blocks_causvid(W, num_layers, x, embed0, grid_sizes, freqs, context, num_heads, kv_caches, kv_start, kv_end, attn="torch_sdpa"):
    """_infer_without_offload - causvid/transformer_infer.py:77-93."""
    for i in range(num_layers):
        x = infer_block_causvid(W, i, x, embed0, grid_sizes, freqs, context, num_heads, kv_caches[i], kv_start, kv_end, attn)
    return x


# ---------------------------------------------------------------------------------------------------------------
# synthetic weights at checkpoint key names (SURVEY.md §8d recipe)
# ---------------------------------------------------------------------------------------------------------------
def synth_block_weights(num_layers: int, dim: int, ffn_dim: int, task: str = "t2v", seed: int = 42, device="cpu") -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    W: Dict[str, torch.Tensor] = {}

    def lin(name, n, k):
        W[name + ".weight"] = (torch.randn(n, k, generator=g) * 0.02).to(torch.bfloat16).to(device)
        W[name + ".bias"] = (torch.randn(n, generator=g) * 0.02).to(torch.bfloat16).to(device)

    def norm_w(name, n):
        W[name] = (1.0 + torch.randn(n, generator=g) * 0.05).to(torch.bfloat16).to(device)

    for i in range(num_layers):
        p = f"blocks.{i}."
        W[p + "modulation"] = (torch.randn(1, 6, dim, generator=g) * 0.1).to(torch.bfloat16).to(device)
        for nm in ("q", "k", "v", "o"):
            lin(p + "self_attn." + nm, dim, dim)
            lin(p + "cross_attn." + nm, dim, dim)
        norm_w(p + "self_attn.norm_q.weight", dim)
        norm_w(p + "self_attn.norm_k.weight", dim)
        norm_w(p + "cross_attn.norm_q.weight", dim)
        norm_w(p + "cross_attn.norm_k.weight", dim)
        norm_w(p + "norm3.weight", dim)
        W[p + "norm3.bias"] = (torch.randn(dim, generator=g) * 0.02).to(torch.bfloat16).to(device)
        if task == "i2v":
            lin(p + "cross_attn.k_img", dim, dim)
            lin(p + "cross_attn.v_img", dim, dim)
            norm_w(p + "cross_attn.norm_k_img.weight", dim)
        lin(p + "ffn.0", ffn_dim, dim)
        lin(p + "ffn.2", dim, ffn_dim)
    return W


def synth_block_inputs(S: int, dim: int, text_len: int = 512, task: str = "t2v", seed: int = 7, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(S, dim, generator=g).to(torch.bfloat16).to(device)
    embed0 = (torch.randn(6, dim, generator=g) * 0.2).to(torch.bfloat16).to(device)
    ctx_rows = text_len + (257 if task == "i2v" else 0)
    context = torch.randn(ctx_rows, dim, generator=g).to(torch.bfloat16).to(device)
    return x, embed0, context


def block_flops(S: int, D: int, F_: int, Lt: int = 512, i2v: bool = False) -> float:
    """FLOPs of one block."""
    fl = 8 * S * D * D + 4 * S * S * D + 4 * S * D * D + 4 * Lt * D * D + 4 * S * Lt * D + 4 * S * D * F_
    if i2v:
        fl += 4 * 257 * D * D + 4 * S * 257 * D
    return float(fl)
