"""ORACLE (test infrastructure, NOT product code).

Plain-torch restatement of CogVideoX's DiT block stack in the reference's DTYPE=BF16 mode
(lightx2v/models/networks/cogvideox/infer/transformer_infer.py:45-145; weights lightx2v/models/networks/cogvideox/weights/
transformers_weights.py:30-77), each function citing the lines it follows.  Pinned: `oracle/gen_golden.py:gen_cogvideox_fixture` runs the
REAL `CogvideoxTransformerInfer` + `CogVideoXBlock` classes on seeded synthetic weights and stores inputs/outputs under
tests/golden/cogvideox_2blocks.safetensors; tests/test_oracle_golden.py requires this restatement to reproduce it bit for bit.
Only tests/ may import this file."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


def _mm(W, name, x):
    """MMWeight.apply = torch.addmm(bias, x, W.t()) - common/ops/mm/mm_weight.py:81-88."""
    return torch.addmm(W[name + ".bias"], x, W[name + ".weight"].t())


def _ln(W, name, x, eps):
    """LNWeight.apply (BF16 mode) - common/ops/norm/layer_norm_weight.py:110."""
    return F.layer_norm(x, (x.shape[-1],), W[name + ".weight"], W[name + ".bias"], eps)


def apply_rotary_emb(x, freqs_cis):
    """transformer_infer.py:5-36 (use_real=True, use_real_unbind_dim=-1): x [H, S, D]; cos/sin [S, D]; fp32 math, one rounding."""
    cos, sin = freqs_cis
    cos, sin = cos[None].to(x.device), sin[None].to(x.device)
    x_real, x_imag = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    x_rotated = torch.stack([-x_imag, x_real], dim=-1).flatten(2)
    return (x.float() * cos + x_rotated.float() * sin).to(x.dtype)


def norm_mod(W, pre, which, hidden, enc, temb):
    """cogvideox_norm1 / cogvideox_norm2 - transformer_infer.py:67-81."""
    t = _mm(W, f"{pre}{which}.linear", F.silu(temb))
    shift, scale, gate, enc_shift, enc_scale, enc_gate = t.chunk(6, dim=1)
    h = _ln(W, f"{pre}{which}.norm", hidden, 1e-5) * (1 + scale)[:, :] + shift[:, :]
    e = _ln(W, f"{pre}{which}.norm", enc, 1e-5) * (1 + enc_scale)[:, :] + enc_shift[:, :]
    return h, e, gate, enc_gate


def attention(W, pre, hidden, enc, rotary, heads):
    """cogvideox_attention - transformer_infer.py:83-113: joint [text ; video] tokens, per-head LayerNorm on q/k, RoPE on the video rows."""
    Lt = enc.size(0)
    x = torch.cat([enc, hidden], dim=0)
    q, k, v = _mm(W, pre + "attn1.to_q", x), _mm(W, pre + "attn1.to_k", x), _mm(W, pre + "attn1.to_v", x)
    d = k.shape[-1] // heads
    q, k, v = (t.view(-1, heads, d).transpose(0, 1) for t in (q, k, v))
    q = _ln(W, pre + "attn1.norm_q", q, 1e-6)
    k = _ln(W, pre + "attn1.norm_k", k, 1e-6)
    q[:, Lt:] = apply_rotary_emb(q[:, Lt:], rotary)
    k[:, Lt:] = apply_rotary_emb(k[:, Lt:], rotary)
    o = F.scaled_dot_product_attention(q[None], k[None], v[None], attn_mask=None, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(1, -1, heads * d).squeeze(0)
    o = _mm(W, pre + "attn1.to_out.0", o)
    e, h = o.split([Lt, o.size(0) - Lt], dim=0)
    return h, e


def infer_block(W, i, hidden, enc, temb, rotary, heads):
    """infer_block - transformer_infer.py:121-145."""
    pre = f"transformer_blocks.{i}."
    Lt = enc.size(0)
    nh, ne, gate, enc_gate = norm_mod(W, pre, "norm1", hidden, enc, temb)
    ah, ae = attention(W, pre, nh, ne, rotary, heads)
    hidden = hidden + gate * ah
    enc = enc + enc_gate * ae
    nh, ne, gate_ff, enc_gate_ff = norm_mod(W, pre, "norm2", hidden, enc, temb)
    x = torch.cat([ne, nh], dim=0)
    ff = _mm(W, pre + "ff.net.2", F.gelu(_mm(W, pre + "ff.net.0.proj", x), approximate="tanh"))     # cogvideox_ff :115-119
    hidden = hidden + gate_ff * ff[Lt:,]
    enc = enc + enc_gate_ff * ff[:Lt,]
    return hidden, enc


def infer_blocks(W, layers, hidden, enc, temb, rotary, heads):
    """CogvideoxTransformerInfer.infer - transformer_infer.py:53-65."""
    for i in range(layers):
        hidden, enc = infer_block(W, i, hidden, enc, temb, rotary, heads)
    return hidden, enc


def rotary_table(frames: int, h: int, w: int, head_dim: int = 64, seed: int = 0):
    """A (cos, sin) pair of the shape and structure get_3d_rotary_pos_embed produces ([S, head_dim] fp32, each angle repeated for the
    two members of a pair); the angles are synthetic (diffusers, which owns that function, is not in the image)."""
    g = torch.Generator().manual_seed(seed)
    ang = torch.rand(frames * h * w, head_dim // 2, generator=g) * 6.2831853
    return ang.cos().repeat_interleave(2, dim=1).contiguous(), ang.sin().repeat_interleave(2, dim=1).contiguous()


def synth_weights(layers: int, dim: int, ff: int, head_dim: int = 64, time_dim: int = 512, seed: int = 42, device="cpu") -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    W: Dict[str, torch.Tensor] = {}

    def lin(name, n, k, scale=0.02):
        W[name + ".weight"] = (torch.randn(n, k, generator=g) * scale).to(torch.bfloat16).to(device)
        W[name + ".bias"] = (torch.randn(n, generator=g) * 0.02).to(torch.bfloat16).to(device)

    def ln(name, n):
        W[name + ".weight"] = (1.0 + torch.randn(n, generator=g) * 0.05).to(torch.bfloat16).to(device)
        W[name + ".bias"] = (torch.randn(n, generator=g) * 0.02).to(torch.bfloat16).to(device)

    for i in range(layers):
        p = f"transformer_blocks.{i}."
        for nm in ("to_q", "to_k", "to_v", "to_out.0"):
            lin(p + "attn1." + nm, dim, dim)
        lin(p + "ff.net.0.proj", ff, dim)
        lin(p + "ff.net.2", dim, ff)
        lin(p + "norm1.linear", 6 * dim, time_dim, 0.01)
        lin(p + "norm2.linear", 6 * dim, time_dim, 0.01)
        ln(p + "attn1.norm_q", head_dim)
        ln(p + "attn1.norm_k", head_dim)
        ln(p + "norm1.norm", dim)
        ln(p + "norm2.norm", dim)
    return W


def synth_inputs(Lt: int, Li: int, dim: int, time_dim: int = 512, seed: int = 7, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    hidden = torch.randn(Li, dim, generator=g).to(torch.bfloat16).to(device)
    enc = torch.randn(Lt, dim, generator=g).to(torch.bfloat16).to(device)
    temb = torch.randn(1, time_dim, generator=g).to(torch.bfloat16).to(device)
    return hidden, enc, temb
