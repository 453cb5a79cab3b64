"""CogVideoX DiT block stack on the sm_100a kernels, with the reference's method surface:
  CogvideoxTransformerInfer.infer(weights, hidden_states, encoder_hidden_states, temb) -> (hidden_states, encoder_hidden_states)
  set_scheduler / infer_block / cogvideox_norm1 / cogvideox_norm2 / cogvideox_attention / cogvideox_ff
(lightx2v/models/networks/cogvideox/infer/transformer_infer.py:45-145).  Weights: this package's `CogvideoxTransformerWeights`
(same attribute names and checkpoint keys as lightx2v/models/networks/cogvideox/weights/transformers_weights.py:5-77) or the
reference's own tree (only `.weight`, `.bias`, `.eps` of the op objects are read).

B200 schedule per block (the reference issues ~45 torch launches incl. two concatenations, three transposes and 8 elementwise passes):
  the text and video streams live in ONE [Lt + Lv, D] buffer for the whole stack (the reference concatenates and splits them twice per
  block, :85 / :112 / :139); AdaLN: LayerNorm(affine) + (1 + scale), shift in one pass per stream (`b200_ln_modulate`); q, k, v as ONE
  [L, D] x [D, 3D] GEMM on the concatenated weight; per-head LayerNorm(64) + pair RoPE on the video rows in one in-place pass
  (`b200_ln_rope_heads64`); attention over the joint sequence with the head_dim-64 FMHA reading q/k/v as strided views of the QKV buffer
  (the reference: F.scaled_dot_product_attention on transposed copies, :105); `x + gate * to_out(attn)` and `x + gate * ff2(...)` as
  GEMM epilogues (two row ranges: text / video gates); GELU(tanh) as the ff.net.0 epilogue."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from .. import lib
from . import ops  # noqa: F401  (registers the op classes under their keys)
from .registry import LN_WEIGHT_REGISTER, MM_KEY, MM_WEIGHT_REGISTER


class CogVideoXBlock:
    """transformers_weights.py:30-77 (attribute names and checkpoint keys)."""

    _MM = {"attn1_to_q": "attn1.to_q", "attn1_to_k": "attn1.to_k", "attn1_to_v": "attn1.to_v", "attn1_to_out": "attn1.to_out.0",
           "ff_net_0_proj": "ff.net.0.proj", "ff_net_2_proj": "ff.net.2", "norm1_linear": "norm1.linear", "norm2_linear": "norm2.linear"}
    _LN = {"attn1_norm_q": ("attn1.norm_q", 1e-6), "attn1_norm_k": ("attn1.norm_k", 1e-6), "norm1_norm": ("norm1.norm", 1e-5), "norm2_norm": ("norm2.norm", 1e-5)}

    def __init__(self, block_index, task="t2v", mm_type=MM_KEY):
        self.block_index, self.task, self.mm_type = block_index, task, mm_type
        self.weight_list = []

    def load_weights(self, weight_dict):
        p = f"transformer_blocks.{self.block_index}."
        for attr, key in self._MM.items():
            setattr(self, attr, MM_WEIGHT_REGISTER[self.mm_type](p + key + ".weight", p + key + ".bias"))
        for attr, (key, eps) in self._LN.items():
            setattr(self, attr, LN_WEIGHT_REGISTER[self.mm_type](p + key + ".weight", p + key + ".bias", eps=eps))
        self.weight_list = [getattr(self, a) for a in list(self._MM) + list(self._LN)]
        for w in self.weight_list:
            w.load(weight_dict)

    def to_cpu(self):
        for w in self.weight_list:
            w.to_cpu()

    def to_cuda(self):
        for w in self.weight_list:
            w.to_cuda()


class CogvideoxTransformerWeights:
    """transformers_weights.py:5-27."""

    def __init__(self, config, task="t2v", mm_type=MM_KEY):
        self.config, self.task, self.mm_type = config, task, mm_type
        self.num_layers = config["num_layers"]
        self.blocks_weights = []

    def load_weights(self, weight_dict):
        self.blocks_weights = [CogVideoXBlock(i, self.task, self.mm_type) for i in range(self.num_layers)]
        for block in self.blocks_weights:
            block.load_weights(weight_dict)

    def to_cpu(self):
        for b in self.blocks_weights:
            b.to_cpu()

    def to_cuda(self):
        for b in self.blocks_weights:
            b.to_cuda()


def rotary_pairs(image_rotary_emb: Tuple[torch.Tensor, torch.Tensor]) -> torch.Tensor:
    """(cos, sin) [S, 64] fp32 with each angle repeated for the two members of a pair (get_3d_rotary_pos_embed's layout, as consumed by
    apply_rotary_emb, transformer_infer.py:21-34) -> [S, 32, 2] fp32 per pair."""
    cos, sin = image_rotary_emb
    if not (torch.equal(cos[:, 0::2], cos[:, 1::2]) and torch.equal(sin[:, 0::2], sin[:, 1::2])):
        raise lib.B200Error("CogvideoxTransformerInfer(B200): rotary table is not pair-structured (cos[2i] != cos[2i+1])")
    return torch.stack([cos[:, 0::2].float(), sin[:, 0::2].float()], dim=-1).contiguous()


class CogvideoxTransformerInfer:
    def __init__(self, config):
        self.config = config
        self.attn_type = "b200_fmha"
        self.num_layers = config["transformer_num_layers"]
        self.heads = config["transformer_num_attention_heads"]
        self.head_dim = config.get("transformer_attention_head_dim", 64)
        if self.head_dim != 64:
            raise lib.B200Error(f"CogvideoxTransformerInfer(B200): head_dim must be 64, got {self.head_dim}")
        self.scheduler = None
        self._bufs: Dict[Tuple, torch.Tensor] = {}
        self._qkv: Dict[int, Tuple] = {}
        self._rope = None

    # ------------------------------------------------------------------ reference surface
    def set_scheduler(self, scheduler):
        self.scheduler = scheduler

    @torch.no_grad()
    def infer(self, weights, hidden_states, encoder_hidden_states, temb):
        Lt, Lv, D = encoder_hidden_states.shape[0], hidden_states.shape[0], hidden_states.shape[1]
        x = self._buf("x", (Lt + Lv, D), hidden_states.device)          # joint residual stream [text ; video], updated in place
        x[:Lt].copy_(encoder_hidden_states)
        x[Lt:].copy_(hidden_states)
        rot = self.scheduler.image_rotary_emb
        for i in range(self.num_layers):
            self._block(weights.blocks_weights[i], x, Lt, temb, rot)
        return x[Lt:].clone(), x[:Lt].clone()

    @torch.no_grad()
    def infer_block(self, weights, hidden_states, encoder_hidden_states, temb, image_rotary_emb):
        """transformer_infer.py:121-145 (one block, separate streams in and out like the reference)."""
        Lt = encoder_hidden_states.shape[0]
        x = torch.cat([encoder_hidden_states, hidden_states], dim=0)
        self._block(weights, x, Lt, temb, image_rotary_emb)
        return x[Lt:], x[:Lt]

    # ------------------------------------------------------------------ helpers
    def _buf(self, name, shape, device):
        key = (name, tuple(shape), str(device))
        b = self._bufs.get(key)
        if b is None:
            b = torch.empty(shape, dtype=torch.bfloat16, device=device)
            self._bufs[key] = b
        return b

    @staticmethod
    def _nk(mm):
        w = mm.weight.t()
        return w if w.is_contiguous() else w.contiguous()

    def _cs(self, rot, device):
        if rot is None:
            return None
        key = (rot[0].data_ptr(), rot[0]._version, tuple(rot[0].shape))
        if self._rope is None or self._rope[0] != key:
            self._rope = (key, rotary_pairs(rot).to(device), rot)
        return self._rope[1]

    def _wqkv(self, w):
        ent = self._qkv.get(id(w))
        ptrs = tuple(m.weight.data_ptr() for m in (w.attn1_to_q, w.attn1_to_k, w.attn1_to_v))
        if ent is None or ent[0] is not w or ent[1] != ptrs:
            wq = torch.cat([self._nk(w.attn1_to_q), self._nk(w.attn1_to_k), self._nk(w.attn1_to_v)], dim=0).contiguous()
            bq = torch.cat([w.attn1_to_q.bias, w.attn1_to_k.bias, w.attn1_to_v.bias]).contiguous()
            ent = (w, ptrs, wq, bq)
            self._qkv[id(w)] = ent
        return ent[2], ent[3]

    def _modulation(self, lin, temb):
        """temb -> silu -> Linear -> six [D] vectors: shift, scale, gate, enc_shift, enc_scale, enc_gate (:68-70)."""
        t = lib.gemm_bf16(F.silu(temb).contiguous(), self._nk(lin), lin.bias)
        return [v.contiguous().reshape(-1) for v in t.chunk(6, dim=1)]

    def _norm(self, ln, lin, x, Lt, temb, out):
        shift, scale, gate, enc_shift, enc_scale, enc_gate = self._modulation(lin, temb)
        lib.ln_modulate(x[:Lt], weight=ln.weight, bias=ln.bias, scale=enc_scale, shift=enc_shift, eps=ln.eps, out=out[:Lt])
        lib.ln_modulate(x[Lt:], weight=ln.weight, bias=ln.bias, scale=scale, shift=shift, eps=ln.eps, out=out[Lt:])
        return gate, enc_gate

    # ------------------------------------------------------------------ one block on the joint stream
    def _block(self, w, x, Lt, temb, rot):
        L, D = x.shape
        H, dev = self.heads, x.device
        n = self._buf("n", (L, D), dev)
        gate, enc_gate = self._norm(w.norm1_norm, w.norm1_linear, x, Lt, temb, n)                       # cogvideox_norm1 :67-73
        wqkv, bqkv = self._wqkv(w)
        qkv = self._buf("qkv", (L, 3 * D), dev)
        lib.gemm_bf16(n, wqkv, bqkv, out=qkv)                                                           # :87-89
        q3 = qkv.view(L, 3, H, 64)
        lib.ln_rope_heads64_(q3[:, 0], w.attn1_norm_q.weight, w.attn1_norm_q.bias, q3[:, 1], w.attn1_norm_k.weight, w.attn1_norm_k.bias,
                             eps=w.attn1_norm_q.eps, cos_sin=self._cs(rot, dev), rope_start=Lt)        # :99-103
        attn = n.view(L, H, 64)                                                                         # the LN scratch is dead: reuse it
        lib.fmha(q3[:, 0], q3[:, 1], q3[:, 2], out=attn)                                                # :105
        wo = self._nk(w.attn1_to_out)
        lib.gemm_bf16(n[:Lt], wo, w.attn1_to_out.bias, out=x[:Lt], epilogue=lib.EPI_GATE_RESIDUAL, gate=enc_gate)   # :132
        lib.gemm_bf16(n[Lt:], wo, w.attn1_to_out.bias, out=x[Lt:], epilogue=lib.EPI_GATE_RESIDUAL, gate=gate)       # :131
        gate_ff, enc_gate_ff = self._norm(w.norm2_norm, w.norm2_linear, x, Lt, temb, n)                 # cogvideox_norm2 :75-81
        w0 = self._nk(w.ff_net_0_proj)
        h = lib.gemm_bf16(n, w0, w.ff_net_0_proj.bias, out=self._buf("h", (L, w0.shape[0]), dev), epilogue=lib.EPI_BIAS_GELU)   # :116-117
        w2 = self._nk(w.ff_net_2_proj)
        lib.gemm_bf16(h[:Lt], w2, w.ff_net_2_proj.bias, out=x[:Lt], epilogue=lib.EPI_GATE_RESIDUAL, gate=enc_gate_ff)   # :143
        lib.gemm_bf16(h[Lt:], w2, w.ff_net_2_proj.bias, out=x[Lt:], epilogue=lib.EPI_GATE_RESIDUAL, gate=gate_ff)       # :142

    # ------------------------------------------------------------------ the reference's sub-steps, for callers that use them directly
    def cogvideox_norm1(self, weights, hidden_states, encoder_hidden_states, temb):
        return self._norm_pair(weights.norm1_norm, weights.norm1_linear, hidden_states, encoder_hidden_states, temb)

    def cogvideox_norm2(self, weights, hidden_states, encoder_hidden_states, temb):
        return self._norm_pair(weights.norm2_norm, weights.norm2_linear, hidden_states, encoder_hidden_states, temb)

    def _norm_pair(self, ln, lin, hidden, enc, temb):
        Lt = enc.shape[0]
        x = torch.cat([enc, hidden], dim=0)
        out = torch.empty_like(x)
        gate, enc_gate = self._norm(ln, lin, x, Lt, temb, out)
        return out[Lt:], out[:Lt], gate.view(1, -1), enc_gate.view(1, -1)

    def cogvideox_ff(self, weights, hidden_states):
        h = lib.gemm_bf16(hidden_states.contiguous(), self._nk(weights.ff_net_0_proj), weights.ff_net_0_proj.bias, epilogue=lib.EPI_BIAS_GELU)
        return lib.gemm_bf16(h, self._nk(weights.ff_net_2_proj), weights.ff_net_2_proj.bias)

    def cogvideox_attention(self, weights, hidden_states, encoder_hidden_states, image_rotary_emb):
        Lt = encoder_hidden_states.shape[0]
        x = torch.cat([encoder_hidden_states, hidden_states], dim=0)
        L, D = x.shape
        wqkv, bqkv = self._wqkv(weights)
        qkv = lib.gemm_bf16(x, wqkv, bqkv)
        q3 = qkv.view(L, 3, self.heads, 64)
        lib.ln_rope_heads64_(q3[:, 0], weights.attn1_norm_q.weight, weights.attn1_norm_q.bias, q3[:, 1], weights.attn1_norm_k.weight, weights.attn1_norm_k.bias,
                             eps=weights.attn1_norm_q.eps, cos_sin=self._cs(image_rotary_emb, x.device), rope_start=Lt)
        o = lib.fmha(q3[:, 0], q3[:, 1], q3[:, 2]).reshape(L, D)
        o = lib.gemm_bf16(o, self._nk(weights.attn1_to_out), weights.attn1_to_out.bias)
        return o[Lt:], o[:Lt]
