"""ORACLE (test infrastructure, NOT product code): plain-torch restatement of the HunyuanVideo DiT blocks as LightX2V runs them
(lightx2v/models/networks/hunyuan/infer/transformer_infer.py: double block :86-277 + :279-306, single block :308-384;
rotary / rms helpers lightx2v/models/networks/hunyuan/infer/utils_bf16.py:5-31), t2v path (token_replace_vec = None).
Weights: flat dict with the checkpoint key names of hunyuan/weights/transformer_weights.py:16-71.
Pinned by tests/golden/hunyuan_blocks_small.safetensors, produced by the REAL HunyuanTransformerInfer (oracle/gen_golden.py)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .wan_oracle import attn_apply, mm_named


def rms_norm(x, weight, eps=1e-6):
    """RMSWeightSgl.apply bf16 fallback over the last (head) dim — rms_norm_weight.py:109-113 (same chain as utils_bf16.py:5-8)."""
    x = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    return x * weight


def rotate_half(x, s0, s1):
    x_real, x_imag = x.reshape(s0, s1, -1, 2).unbind(-1)
    return torch.stack([-x_imag, x_real], dim=-1).flatten(2)          # utils_bf16.py:11-13


def apply_rotary_emb(xq, xk, freqs_cis):
    """utils_bf16.py:21-31."""
    s0, s1, s2 = xq.shape
    cos = freqs_cis[0].view(s0, 1, s2)
    sin = freqs_cis[1].view(s0, 1, s2)
    return xq * cos + rotate_half(xq, s0, s1) * sin, xk * cos + rotate_half(xk, s0, s1) * sin


def varlen_attention(q, k, v, cu_seqlens, impl="torch_sdpa"):
    """ATTN op with cu_seqlens = [0, img + txt_valid, img + txt_padded] (pre_infer.py:50-58): two independent segments."""
    bounds = [int(b) for b in cu_seqlens]
    outs = [attn_apply(q[a:b], k[a:b], v[a:b], impl) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
    return torch.cat(outs, dim=0)


def _qkv_heads(qkv, heads):
    L = qkv.shape[0]
    return qkv.view(L, 3, heads, -1).unbind(1)                         # "L (K H D) -> K L H D"


def _mod_rows(h, mod, tr, first, i_scale, i_shift):
    """`norm * (1 + scale) + shift` with the token-replace modulation on the first `first` rows (i2v: the tokens of the conditioning
    frame follow the t = 0 embedding) - transformer_infer.py:204-209, 281-286, 323-328."""
    if tr is None:
        return h * (1 + mod[i_scale]) + mod[i_shift]
    return torch.concat((h[:first] * (1 + tr[i_scale]) + tr[i_shift], h[first:] * (1 + mod[i_scale]) + mod[i_shift]), dim=0)


def _gate_rows(out, mod, tr, first, i_gate):
    """`out * gate` with the token-replace gate on the first rows - transformer_infer.py:192-198, 373-378."""
    if tr is None:
        return out * mod[i_gate]
    return torch.concat((out[:first] * tr[i_gate], out[first:] * mod[i_gate]), dim=0)


def infer_double_block(W: Dict[str, torch.Tensor], i: int, img, txt, vec, cu_seqlens, freqs_cis, heads: int, attn="torch_sdpa", token_replace_vec=None,
                       first_frame_tokens=None):
    """transformer_infer.py:234-277 (phases :86-232).  NOTE: with token replacement the reference's phase 3 (:224-232) applies the ORDINARY
    img_mod2_gate to every image row (tr_img_mod2_gate is computed and returned by phase 1 but never used) - restated as is."""
    p = f"double_blocks.{i}."
    vec_silu = F.silu(vec)
    im = mm_named(W, p + "img_mod.linear", vec_silu).chunk(6, dim=-1)      # shift1, scale1, gate1, shift2, scale2, gate2
    tm = mm_named(W, p + "txt_mod.linear", vec_silu).chunk(6, dim=-1)
    tr = mm_named(W, p + "img_mod.linear", F.silu(token_replace_vec)).chunk(6, dim=-1) if token_replace_vec is not None else None     # :100-103

    def pre_atten(x, mod, pre, rope):
        x_mod = _mod_rows(F.layer_norm(x, (x.shape[1],), None, None, 1e-6), mod, tr if rope else None, first_frame_tokens, 1, 0)   # :280-286
        q, k, v = _qkv_heads(mm_named(W, p + pre + "_attn_qkv", x_mod), heads)
        q = rms_norm(q, W[p + pre + "_attn_q_norm.weight"])
        k = rms_norm(k, W[p + pre + "_attn_k_norm.weight"])
        if rope:
            q, k = apply_rotary_emb(q, k, freqs_cis)
        return q, k, v

    iq, ik, iv = pre_atten(img, im, "img", True)
    tq, tk, tv = pre_atten(txt, tm, "txt", False)
    q, k, v = torch.cat((iq, tq)), torch.cat((ik, tk)), torch.cat((iv, tv))
    attn_out = varlen_attention(q, k, v, cu_seqlens, attn)
    img_attn, txt_attn = attn_out[: img.shape[0]], attn_out[img.shape[0]:]
    img_out = mm_named(W, p + "img_attn_proj", img_attn)
    txt_out = mm_named(W, p + "txt_attn_proj", txt_attn)

    def mlp(x, out, mod, pre):
        t = tr if pre == "img" else None
        x = x + _gate_rows(out, mod, t, first_frame_tokens, 2)                                                 # :192-199 / :215-216
        h = _mod_rows(F.layer_norm(x, (x.shape[1],), None, None, 1e-6), mod, t, first_frame_tokens, 4, 3)
        h = F.gelu(mm_named(W, p + pre + "_mlp.fc1", h), approximate="tanh")
        h = mm_named(W, p + pre + "_mlp.fc2", h)
        return x + h * mod[5]                                                                                  # phase 3 :224-232

    return mlp(img, img_out, im, "img"), mlp(txt, txt_out, tm, "txt")


def infer_single_block(W, i: int, x, vec, txt_seq_len: int, cu_seqlens, freqs_cis, heads: int, hidden: int, attn="torch_sdpa", token_replace_vec=None,
                       first_frame_tokens=None):
    """transformer_infer.py:308-384."""
    p = f"single_blocks.{i}."
    mod = mm_named(W, p + "modulation.linear", F.silu(vec)).chunk(3, dim=-1)                     # shift, scale, gate
    tr = mm_named(W, p + "modulation.linear", F.silu(token_replace_vec)).chunk(3, dim=-1) if token_replace_vec is not None else None
    mod_gate = mod[2]
    x_mod = _mod_rows(F.layer_norm(x, (x.shape[1],), None, None, 1e-6), mod, tr, first_frame_tokens, 1, 0)
    x_mod = mm_named(W, p + "linear1", x_mod)
    qkv, mlp = torch.split(x_mod, [3 * hidden, x_mod.shape[1] - 3 * hidden], dim=-1)
    q, k, v = _qkv_heads(qkv, heads)
    q = rms_norm(q, W[p + "q_norm.weight"])
    k = rms_norm(k, W[p + "k_norm.weight"])
    img_q, txt_q = q[:-txt_seq_len], q[-txt_seq_len:]
    img_k, txt_k = k[:-txt_seq_len], k[-txt_seq_len:]
    img_q, img_k = apply_rotary_emb(img_q, img_k, freqs_cis)
    q, k = torch.cat((img_q, txt_q)), torch.cat((img_k, txt_k))
    attn_out = varlen_attention(q, k, v, cu_seqlens, attn)
    out = torch.cat((attn_out, F.gelu(mlp, approximate="tanh")), 1)
    out = mm_named(W, p + "linear2", out)
    return x + _gate_rows(out, mod, tr, first_frame_tokens, 2)                                                 # :371-379


def infer_blocks(W, n_double: int, n_single: int, img, txt, vec, cu_seqlens, freqs_cis, heads: int, attn="torch_sdpa", token_replace_vec=None,
                 first_frame_tokens=None):
    """_infer_without_offload — transformer_infer.py:71-84."""
    hidden = img.shape[1]
    for i in range(n_double):
        img, txt = infer_double_block(W, i, img, txt, vec, cu_seqlens, freqs_cis, heads, attn, token_replace_vec, first_frame_tokens)
    x = torch.cat((img, txt), 0)
    for i in range(n_single):
        x = infer_single_block(W, i, x, vec, txt.shape[0], cu_seqlens, freqs_cis, heads, hidden, attn, token_replace_vec, first_frame_tokens)
    return x[: img.shape[0]]


def synth_weights(n_double: int, n_single: int, hidden: int, mlp_hidden: int, seed=42, device="cpu") -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    W: Dict[str, torch.Tensor] = {}

    def lin(name, n, k, scale=0.02):
        W[name + ".weight"] = (torch.randn(n, k, generator=g) * scale).to(torch.bfloat16).to(device)
        W[name + ".bias"] = (torch.randn(n, generator=g) * 0.02).to(torch.bfloat16).to(device)

    def norm(name):
        W[name] = (1.0 + torch.randn(128, generator=g) * 0.05).to(torch.bfloat16).to(device)

    for i in range(n_double):
        p = f"double_blocks.{i}."
        for s in ("img", "txt"):
            lin(p + s + "_mod.linear", 6 * hidden, hidden, 0.01)
            lin(p + s + "_attn_qkv", 3 * hidden, hidden)
            norm(p + s + "_attn_q_norm.weight")
            norm(p + s + "_attn_k_norm.weight")
            lin(p + s + "_attn_proj", hidden, hidden)
            lin(p + s + "_mlp.fc1", mlp_hidden, hidden)
            lin(p + s + "_mlp.fc2", hidden, mlp_hidden)
    for i in range(n_single):
        p = f"single_blocks.{i}."
        lin(p + "linear1", 3 * hidden + mlp_hidden, hidden)
        lin(p + "linear2", hidden, hidden + mlp_hidden)
        norm(p + "q_norm.weight")
        norm(p + "k_norm.weight")
        lin(p + "modulation.linear", 3 * hidden, hidden, 0.01)
    return W


def synth_inputs(img_len: int, txt_len: int, txt_valid: int, hidden: int, seed=7, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(img_len, hidden, generator=g).to(torch.bfloat16).to(device)
    txt = torch.randn(txt_len, hidden, generator=g).to(torch.bfloat16).to(device)
    vec = torch.randn(1, hidden, generator=g).to(torch.bfloat16).to(device)
    ang = torch.rand(img_len, 64, generator=g) * 6.28
    cos = ang.cos().repeat_interleave(2, dim=1).to(torch.bfloat16).to(device)      # scheduler.py:58-59, cast :318-319
    sin = ang.sin().repeat_interleave(2, dim=1).to(torch.bfloat16).to(device)
    cu = [0, img_len + txt_valid, img_len + txt_len]
    return img, txt, vec, cu, (cos, sin)


# ---------------------------------------------------------------------------------------------------------------
# pre-infer / post-infer (lightx2v/models/networks/hunyuan/infer/pre_infer.py, post_infer.py), t2v
# Pinned by tests/golden/hunyuan_prepost.safetensors, produced by the REAL HunyuanPreInfer / HunyuanPostInfer methods
# (oracle/gen_golden.py:gen_hunyuan_prepost_fixture).  `infer_text_in` (the two-block token refiner, pre_infer.py:83-138) is NOT restated:
# at this snapshot the reference's own call raises (it hands [1, L, H, D] tensors to TorchSDPAWeight.apply, which unsqueezes again,
# attn_weight.py:229-235), so there is no reference output to pin a restatement to.
# ---------------------------------------------------------------------------------------------------------------
def _timestep_embedding(t: torch.Tensor) -> torch.Tensor:
    """pre_infer.py:69-71 / 84-86 / 146-148: 256-wide sinusoidal embedding, fp32 math, one rounding to bf16."""
    import math
    freqs = torch.exp(-math.log(10000) * torch.arange(start=0, end=128, dtype=torch.float32) / 128).to(device=t.device)
    args = t.float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1).to(dtype=torch.bfloat16)


def infer_time_in(W, t):
    """pre_infer.py:68-75 (t: 0-d tensor)."""
    emb = _timestep_embedding(t.unsqueeze(0).unsqueeze(0))
    return mm_named(W, "time_in.mlp.2", F.silu(mm_named(W, "time_in.mlp.0", emb)))


def infer_guidance_in(W, guidance):
    """pre_infer.py:145-152 (guidance: [1] bf16 tensor = embedded_guidance_scale * 1000)."""
    emb = _timestep_embedding(guidance)
    return mm_named(W, "guidance_in.mlp.2", F.silu(mm_named(W, "guidance_in.mlp.0", emb)))


def infer_vector_in(W, text_states_2):
    """pre_infer.py:139-143."""
    return mm_named(W, "vector_in.out_layer", F.silu(mm_named(W, "vector_in.in_layer", text_states_2)))


def infer_img_in(W, x):
    """pre_infer.py:77-80: Conv3d(16 -> hidden, kernel = stride = (1, 2, 2)) on [1, 16, T, H, W], flattened to [1, T*H/2*W/2, hidden]."""
    out = F.conv3d(x, W["img_in.proj.weight"], W["img_in.proj.bias"], stride=(1, 2, 2))
    return out.flatten(2).transpose(1, 2)


def cu_seqlens(text_mask: torch.Tensor, img_seq_len: int):
    """pre_infer.py:45-58 (host ints)."""
    bs, L = text_mask.shape
    text_len = text_mask.sum(dim=1)
    max_len = L + img_seq_len
    cu = [0] * (2 * bs + 1)
    for i in range(bs):
        cu[2 * i + 1] = i * max_len + int(text_len[i]) + img_seq_len
        cu[2 * i + 2] = (i + 1) * max_len
    return cu


def post_infer(W, img, vec, latent_shape):
    """post_infer.py:11-33: adaLN (shift, scale) from silu(vec), LayerNorm(no affine, eps 1e-6) * (1 + scale) + shift in bf16, the
    final linear in FP32 (MM_WEIGHT "Default-Force-FP32", post_weights.py:10), unpatchify to [1, 16, T, H, W] fp32."""
    out = mm_named(W, "final_layer.adaLN_modulation.1", F.silu(vec))
    shift, scale = out.chunk(2, dim=1)
    out = F.layer_norm(img, (img.shape[1],), None, None, 1e-6)
    out = out * (1 + scale) + shift
    out = torch.addmm(W["final_layer.linear.bias"].float(), out.to(torch.float32), W["final_layer.linear.weight"].float().t())
    _, _, ot, oh, ow = latent_shape
    tt, th, tw = ot, oh // 2, ow // 2
    out = out.reshape(shape=(1, tt, th, tw, 16, 1, 2, 2))
    out = torch.einsum("nthwcopq->nctohpwq", out)
    return out.reshape(shape=(1, 16, tt, th * 2, tw * 2))


def synth_prepost_weights(hidden: int = 3072, seed: int = 11, device="cpu") -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    W: Dict[str, torch.Tensor] = {}

    def lin(name, n, k):
        W[name + ".weight"] = (torch.randn(n, k, generator=g) * k ** -0.5).to(torch.bfloat16).to(device)
        W[name + ".bias"] = (torch.randn(n, generator=g) * 0.02).to(torch.bfloat16).to(device)

    lin("time_in.mlp.0", hidden, 256)
    lin("time_in.mlp.2", hidden, hidden)
    lin("guidance_in.mlp.0", hidden, 256)
    lin("guidance_in.mlp.2", hidden, hidden)
    lin("vector_in.in_layer", hidden, 768)
    lin("vector_in.out_layer", hidden, hidden)
    W["img_in.proj.weight"] = (torch.randn(hidden, 16, 1, 2, 2, generator=g) * 64 ** -0.5).to(torch.bfloat16).to(device)
    W["img_in.proj.bias"] = (torch.randn(hidden, generator=g) * 0.02).to(torch.bfloat16).to(device)
    lin("final_layer.adaLN_modulation.1", 2 * hidden, hidden)
    lin("final_layer.linear", 64, hidden)
    return W
