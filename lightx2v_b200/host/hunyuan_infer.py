"""HunyuanVideo DiT blocks (20 double-stream + 40 single-stream) on the sm_100a kernels, with the reference's method surface:
  HunyuanTransformerInfer.infer(weights, img, txt, vec, cu_seqlens_qkv, max_seqlen_qkv, freqs_cis, token_replace_vec=None, frist_frame_token_num=None)
  infer_double_block / infer_single_block          (lightx2v/models/networks/hunyuan/infer/transformer_infer.py:30-84, 234-277, 381-384)
Weights: this package's tree (`HunyuanTransformerWeights` below, same attribute and checkpoint key names as
lightx2v/models/networks/hunyuan/weights/transformer_weights.py:5-71) or the reference's own tree.

Fusions (per double block ~60 reference launches -> 20): LN + modulate in one pass; the image and text QKV GEMMs write into ONE joint
[img+txt, 3D] buffer so the q/k/v concatenations (:121-123) disappear; per-head q/k RMSNorm + RoPE in one in-place pass;
the two varlen segments of cu_seqlens = [0, img+txt_valid, img+txt_pad] (pre_infer.py:50-58) are two FMHA launches on strided views;
`x + gate * proj(attn)` and `x + gate * fc2(...)` are GEMM epilogues; GELU(tanh) is the fc1 epilogue.
Single block: linear1 is issued as two GEMMs on row-slices of its weight (qkv part, mlp part with GELU epilogue) into an [L, 8D] buffer
laid out [q | k | v | attn | gelu(mlp)] so that linear2 reads its [attn | mlp] operand in place (no torch.cat, :368).
i2v token replacement (token_replace_vec, frist_frame_token_num; :100-103, 192-209, 281-286, 316-328, 373-378): the tokens of the conditioning
frame follow the t = 0 embedding, i.e. the first rows of the image stream take their shift / scale / gate from `mod(silu(token_replace_vec))` -
here simply a second `ln_modulate` / GEMM-epilogue call on that row range (like the reference, phase 3 of the double block gates every image
row with the ordinary gate)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from .. import lib
from .ops import ATTN_KEY, MM_KEY
from .registry import ATTN_WEIGHT_REGISTER, MM_WEIGHT_REGISTER, RMS_WEIGHT_REGISTER
from .weight_module import WeightModule, WeightModuleList


class HunyuanTransformerDoubleBlock(WeightModule):
    def __init__(self, block_index, config):
        super().__init__()
        self.block_index, self.config = block_index, config
        mm = MM_WEIGHT_REGISTER[(config.get("mm_config") or {}).get("mm_type", MM_KEY)]
        p = f"double_blocks.{block_index}."
        for s in ("img", "txt"):
            self.add_module(f"{s}_mod", mm(p + f"{s}_mod.linear.weight", p + f"{s}_mod.linear.bias"))
            self.add_module(f"{s}_attn_qkv", mm(p + f"{s}_attn_qkv.weight", p + f"{s}_attn_qkv.bias"))
            self.add_module(f"{s}_attn_q_norm", RMS_WEIGHT_REGISTER["sgl-kernel"](p + f"{s}_attn_q_norm.weight", eps=1e-6))
            self.add_module(f"{s}_attn_k_norm", RMS_WEIGHT_REGISTER["sgl-kernel"](p + f"{s}_attn_k_norm.weight", eps=1e-6))
            self.add_module(f"{s}_attn_proj", mm(p + f"{s}_attn_proj.weight", p + f"{s}_attn_proj.bias"))
            self.add_module(f"{s}_mlp_fc1", mm(p + f"{s}_mlp.fc1.weight", p + f"{s}_mlp.fc1.bias"))
            self.add_module(f"{s}_mlp_fc2", mm(p + f"{s}_mlp.fc2.weight", p + f"{s}_mlp.fc2.bias"))
        self.add_module("double_attn", ATTN_WEIGHT_REGISTER[config.get("attention_type", ATTN_KEY)]())


class HunyuanTransformerSingleBlock(WeightModule):
    def __init__(self, block_index, config):
        super().__init__()
        self.block_index, self.config = block_index, config
        mm = MM_WEIGHT_REGISTER[(config.get("mm_config") or {}).get("mm_type", MM_KEY)]
        p = f"single_blocks.{block_index}."
        self.add_module("linear1", mm(p + "linear1.weight", p + "linear1.bias"))
        self.add_module("linear2", mm(p + "linear2.weight", p + "linear2.bias"))
        self.add_module("q_norm", RMS_WEIGHT_REGISTER["sgl-kernel"](p + "q_norm.weight", eps=1e-6))
        self.add_module("k_norm", RMS_WEIGHT_REGISTER["sgl-kernel"](p + "k_norm.weight", eps=1e-6))
        self.add_module("modulation", mm(p + "modulation.linear.weight", p + "modulation.linear.bias"))
        self.add_module("single_attn", ATTN_WEIGHT_REGISTER[config.get("attention_type", ATTN_KEY)]())


class HunyuanTransformerWeights(WeightModule):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.double_blocks_num = config.get("double_blocks_num", 20)
        self.single_blocks_num = config.get("single_blocks_num", 40)
        self.add_module("double_blocks", WeightModuleList([HunyuanTransformerDoubleBlock(i, config) for i in range(self.double_blocks_num)]))
        self.add_module("single_blocks", WeightModuleList([HunyuanTransformerSingleBlock(i, config) for i in range(self.single_blocks_num)]))


def rope_cos_sin_pairs(freqs_cis: Tuple[torch.Tensor, torch.Tensor]) -> torch.Tensor:
    """(cos, sin) [S, 128] with repeat-interleaved pairs (schedulers/hunyuan/scheduler.py:58-59) -> [S, 64, 2] fp32 per pair."""
    cos, sin = freqs_cis
    return torch.stack([cos[:, 0::2].float(), sin[:, 0::2].float()], dim=-1).contiguous()


class HunyuanTransformerInfer:
    def __init__(self, config):
        self.config = config
        self.attention_type = config.get("attention_type", ATTN_KEY)
        self.double_blocks_num = config.get("double_blocks_num", 20)
        self.single_blocks_num = config.get("single_blocks_num", 40)
        self.heads_num = config.get("heads_num", 24)
        self.hidden_size = config.get("hidden_size", 3072)
        self.mlp_hidden_dim = config.get("mlp_hidden_dim", 12288)
        if self.hidden_size // self.heads_num != 128:
            raise lib.B200Error("HunyuanTransformerInfer(B200): head_dim must be 128")
        self.parallel_attention = None
        self._bufs: Dict[Tuple, torch.Tensor] = {}
        self._rope: Dict[Tuple, torch.Tensor] = {}
        self.infer_func = self._infer_without_offload

    # ------------------------------------------------------------------ reference surface
    def set_scheduler(self, scheduler):
        self.scheduler = scheduler

    def infer(self, weights, img, txt, vec, cu_seqlens_qkv, max_seqlen_qkv, freqs_cis, token_replace_vec=None, frist_frame_token_num=None):
        if token_replace_vec is not None and self.parallel_attention is not None:
            raise lib.B200Error("HunyuanTransformerInfer(B200): token replacement together with sequence parallelism is not implemented")
        return self.infer_func(weights, img, txt, vec, cu_seqlens_qkv, max_seqlen_qkv, freqs_cis, token_replace_vec, frist_frame_token_num)

    def _infer_without_offload(self, weights, img, txt, vec, cu_seqlens_qkv, max_seqlen_qkv, freqs_cis, token_replace_vec, frist_frame_token_num):
        img_seq_len = img.shape[0]
        for i in range(self.double_blocks_num):
            img, txt = self.infer_double_block(weights.double_blocks[i], img, txt, vec, cu_seqlens_qkv, max_seqlen_qkv, freqs_cis, token_replace_vec,
                                               frist_frame_token_num)
        x = torch.cat((img, txt), 0)
        for i in range(self.single_blocks_num):
            x = self.infer_single_block(weights.single_blocks[i], x, vec, txt.shape[0], cu_seqlens_qkv, max_seqlen_qkv, freqs_cis, token_replace_vec,
                                        frist_frame_token_num)
        return x[:img_seq_len, ...], vec

    # ------------------------------------------------------------------ helpers
    def _buf(self, name, shape, device):
        key = (name, tuple(shape), str(device))
        b = self._bufs.get(key)
        if b is None:
            b = torch.empty(shape, dtype=torch.bfloat16, device=device)
            self._bufs[key] = b
        return b

    def _cs(self, freqs_cis):
        key = (freqs_cis[0].data_ptr(), freqs_cis[0]._version, tuple(freqs_cis[0].shape))
        t = self._rope.get(key)
        if t is None:
            t = rope_cos_sin_pairs(freqs_cis)
            self._rope = {key: t}
        return t

    @staticmethod
    def _bounds(cu):
        return [int(v) for v in (cu.tolist() if isinstance(cu, torch.Tensor) else cu)]

    @staticmethod
    def _nk(mm):
        w = mm.weight.t()
        return w if w.is_contiguous() else w.contiguous()

    @staticmethod
    def _ln_rows(x, mod, tr, first, i_scale, i_shift, out):
        """LayerNorm (no affine) + modulation into `out`; with token replacement the first `first` rows use the `tr` vectors."""
        if tr is None or not first:
            return lib.ln_modulate(x, scale=mod[i_scale], shift=mod[i_shift], out=out)
        lib.ln_modulate(x[:first], scale=tr[i_scale], shift=tr[i_shift], out=out[:first])
        lib.ln_modulate(x[first:], scale=mod[i_scale], shift=mod[i_shift], out=out[first:])
        return out

    def _gated_linear(self, a, mm, x, mod, tr, first, i_gate):
        """x += gate * (a @ W^T + b) as a GEMM epilogue; two row ranges when the first rows carry the token-replace gate."""
        w = self._nk(mm)
        if tr is None or not first:
            return lib.gemm_bf16(a, w, mm.bias, out=x, epilogue=lib.EPI_GATE_RESIDUAL, gate=mod[i_gate])
        lib.gemm_bf16(a[:first], w, mm.bias, out=x[:first], epilogue=lib.EPI_GATE_RESIDUAL, gate=tr[i_gate])
        lib.gemm_bf16(a[first:], w, mm.bias, out=x[first:], epilogue=lib.EPI_GATE_RESIDUAL, gate=mod[i_gate])
        return x

    def _attention(self, qkv3, bounds, out, o_cols, txt_len=None):
        """qkv3: [L, 3, H, 128] view; out: [L, *] buffer whose columns o_cols hold the attention output [L, H*128].
        With `parallel_attention` set (Ulysses, host/ulysses.py:HunyuanUlyssesAttention) the image rows are this rank's shard,
        the text rows are replicated and `bounds` are the GLOBAL cu_seqlens [0, img_total + txt_valid, img_total + txt_len]."""
        H = self.heads_num
        if self.parallel_attention is not None:
            L = qkv3.shape[0]
            self.parallel_attention(qkv3, L - txt_len, bounds, out[:, o_cols[0]:o_cols[1]].unflatten(1, (H, 128)))
            return
        for a, b in zip(bounds[:-1], bounds[1:]):
            if b > a:
                o = out[a:b, o_cols[0]:o_cols[1]].unflatten(1, (H, 128))
                lib.fmha(qkv3[a:b, 0], qkv3[a:b, 1], qkv3[a:b, 2], out=o)

    # ------------------------------------------------------------------ double-stream block
    def infer_double_block(self, weights, img, txt, vec, cu_seqlens_qkv, max_seqlen_qkv, freqs_cis, token_replace_vec=None, frist_frame_token_num=None):
        D, H = self.hidden_size, self.heads_num
        Li, Lt = img.shape[0], txt.shape[0]
        L = Li + Lt
        dev = img.device
        vec_silu = F.silu(vec)
        im = weights.img_mod.apply(vec_silu).reshape(6, D)            # shift1, scale1, gate1, shift2, scale2, gate2   (:89-97)
        tm = weights.txt_mod.apply(vec_silu).reshape(6, D)
        tr = weights.img_mod.apply(F.silu(token_replace_vec)).reshape(6, D) if token_replace_vec is not None else None     # :100-103
        first = int(frist_frame_token_num) if tr is not None else 0
        qkv = self._buf("qkv", (L, 3 * D), dev)                       # joint [img; txt] buffer: no torch.cat of q, k, v
        n_img = self._ln_rows(img, im, tr, first, 1, 0, self._buf("n_img", (Li, D), dev))
        lib.gemm_bf16(n_img, self._nk(weights.img_attn_qkv), weights.img_attn_qkv.bias, out=qkv[:Li])
        n_txt = lib.ln_modulate(txt, scale=tm[1], shift=tm[0], out=self._buf("n_txt", (Lt, D), dev))
        lib.gemm_bf16(n_txt, self._nk(weights.txt_attn_qkv), weights.txt_attn_qkv.bias, out=qkv[Li:])
        q3 = qkv.view(L, 3, H, 128)
        cs = self._cs(freqs_cis)
        lib.rms_rope_heads_(q3[:Li, 0], weights.img_attn_q_norm.weight, q3[:Li, 1], weights.img_attn_k_norm.weight,
                            eps=weights.img_attn_q_norm.eps, cos_sin=cs, rope_rows=Li)
        lib.rms_rope_heads_(q3[Li:, 0], weights.txt_attn_q_norm.weight, q3[Li:, 1], weights.txt_attn_k_norm.weight,
                            eps=weights.txt_attn_q_norm.eps)
        attn = self._buf("attn", (L, D), dev)
        self._attention(q3, self._bounds(cu_seqlens_qkv), attn, (0, D), txt_len=Lt)
        # x = x + proj(attn) * gate1  -> GEMM epilogue, in place on the stream tensors (the reference makes new tensors, same values)
        self._gated_linear(attn[:Li], weights.img_attn_proj, img, im, tr, first, 2)
        lib.gemm_bf16(attn[Li:], self._nk(weights.txt_attn_proj), weights.txt_attn_proj.bias, out=txt, epilogue=lib.EPI_GATE_RESIDUAL, gate=tm[2])
        for x, mod, fc1, fc2, nm in ((img, im, weights.img_mlp_fc1, weights.img_mlp_fc2, "img"), (txt, tm, weights.txt_mlp_fc1, weights.txt_mlp_fc2, "txt")):
            n = self._ln_rows(x, mod, tr if nm == "img" else None, first, 4, 3, self._buf("n_" + nm, tuple(x.shape), dev))
            h = lib.gemm_bf16(n, self._nk(fc1), fc1.bias, out=self._buf("h_" + nm, (x.shape[0], self.mlp_hidden_dim), dev), epilogue=lib.EPI_BIAS_GELU)
            lib.gemm_bf16(h, self._nk(fc2), fc2.bias, out=x, epilogue=lib.EPI_GATE_RESIDUAL, gate=mod[5])
        return img, txt

    # ------------------------------------------------------------------ single-stream block
    def infer_single_block(self, weights, x, vec, txt_seq_len, cu_seqlens_qkv, max_seqlen_qkv, freqs_cis, token_replace_vec=None, frist_frame_token_num=None):
        D, H, M = self.hidden_size, self.heads_num, self.mlp_hidden_dim
        L = x.shape[0]
        dev = x.device
        mod = weights.modulation.apply(F.silu(vec)).reshape(3, D)     # shift, scale, gate   (:309-311)
        tr = weights.modulation.apply(F.silu(token_replace_vec)).reshape(3, D) if token_replace_vec is not None else None    # :313-316
        first = int(frist_frame_token_num) if tr is not None else 0
        n = self._ln_rows(x, mod, tr, first, 1, 0, self._buf("n_x", (L, D), dev))
        w1, b1 = self._nk(weights.linear1), weights.linear1.bias
        buf = self._buf("lin1", (L, 4 * D + M), dev)                  # [q | k | v | attn | gelu(mlp)]
        lib.gemm_bf16(n, w1[: 3 * D], b1[: 3 * D], out=buf[:, : 3 * D])
        lib.gemm_bf16(n, w1[3 * D:], b1[3 * D:], out=buf[:, 4 * D:], epilogue=lib.EPI_BIAS_GELU)
        q3 = buf[:, : 3 * D].unflatten(1, (3, H, 128))
        Li = L - txt_seq_len
        cs = self._cs(freqs_cis)
        lib.rms_rope_heads_(q3[:, 0], weights.q_norm.weight, q3[:, 1], weights.k_norm.weight, eps=weights.q_norm.eps, cos_sin=cs, rope_rows=Li)
        self._attention(q3, self._bounds(cu_seqlens_qkv), buf, (3 * D, 4 * D), txt_len=txt_seq_len)
        # x = x + linear2([attn | gelu(mlp)]) * gate
        self._gated_linear(buf[:, 3 * D:], weights.linear2, x, mod, tr, first, 2)
        return x
