"""CausVid block variant on the B200 kernels: autoregressive generation chunk by chunk, self-attention of the current chunk against a K/V
cache of all previous chunks.

Mirrors WanTransformerInferCausVid (lightx2v/models/networks/wan/infer/causvid/transformer_infer.py:8-220): same constructor keys
(`num_frames`, `num_frame_per_block`, `frame_seq_length`, `text_len`), same `infer(..., kv_start, kv_end)` signature, K/V caches of
`num_frames * frame_seq_length` rows per block.  B200 mapping of infer_self_attn (:95-137):
  * the K and V projections write STRAIGHT into the cache rows [kv_start, kv_end) (GEMM output pointer = cache slice; the reference
    computes k, v and copies them in, :112-113);
  * RMSNorm + RoPE run in place on q and on that cache slice, with the RoPE table indexed at `row + kv_start` (compute_freqs_causvid's
    start_frame offset, utils.py:62-75) - one kernel;
  * the FMHA reads q [S, H, 128] and the cache prefix [:kv_end] in place (sq != sk), the o-projection epilogue applies x += gate * y.
Cross-attention (prompt K/V cached after the first chunk, :146-155) and the FFN are the base class's fused paths."""
from __future__ import annotations

import torch

from .. import lib
from .wan_infer import WanTransformerInfer


class WanTransformerInferCausVid(WanTransformerInfer):
    def __init__(self, config):
        super().__init__(config)
        self.num_frames = config["num_frames"]
        self.num_frame_per_block = config["num_frame_per_block"]
        self.frame_seq_length = config["frame_seq_length"]
        self.text_len = config["text_len"]
        self.kv_cache = None
        self.crossattn_cache = None
        self.native_block = False                   # the native block driver has no KV-cache variant

    def _init_kv_cache(self, dtype, device):
        kv_size = self.num_frames * self.frame_seq_length
        self.kv_cache = [{"k": torch.zeros([kv_size, self.num_heads, self.head_dim], dtype=dtype, device=device),
                          "v": torch.zeros([kv_size, self.num_heads, self.head_dim], dtype=dtype, device=device)} for _ in range(self.blocks_num)]

    def _init_crossattn_cache(self, dtype, device):
        """Kept for interface parity (:32-45): the prompt K/V cache of the base class is keyed on the context tensor and needs no allocation."""
        self.crossattn_cache = [{"is_init": False} for _ in range(self.blocks_num)]
        for c in self._caches.values():
            c.kv.clear()

    def infer(self, weights, grid_sizes, embed, x, embed0, seq_lens, freqs, context, kv_start, kv_end):
        if self.kv_cache is None:
            self._init_kv_cache(torch.bfloat16, x.device)
        for block_idx in range(self.blocks_num):
            x = self.infer_block(weights.blocks[block_idx], grid_sizes, embed, x, embed0, seq_lens, freqs, context, block_idx, kv_start, kv_end)
        return x

    def infer_block(self, weights, grid_sizes, embed, x, embed0, seq_lens, freqs, context, block_idx=None, kv_start=None, kv_end=None):
        shift_msa, scale_msa, gate_msa, c_shift_msa, c_scale_msa, c_gate_msa = self.infer_modulation(weights.compute_phases[0], embed0)
        x = self.infer_self_attn_cached(weights.compute_phases[1], grid_sizes, x, freqs, shift_msa, scale_msa, gate_msa, block_idx, kv_start, kv_end)
        x, attn_out = self.infer_cross_attn(weights.compute_phases[2], x, context, None, None)
        y = self.infer_ffn(weights.compute_phases[3], x, attn_out, c_shift_msa, c_scale_msa, c_gate_msa=c_gate_msa)
        return self.post_process(x, y, c_gate_msa)

    def infer_self_attn_cached(self, weights, grid_sizes, x, freqs, shift_msa, scale_msa, gate_msa, block_idx, kv_start, kv_end):
        if self.parallel_attention is not None:
            raise NotImplementedError("Parallel attention is not implemented for causvid inference")     # as the reference, :106-107
        S, D = x.shape
        H, d = self.num_heads, self.head_dim
        if kv_end - kv_start != S:
            raise lib.B200Error(f"causvid: chunk of {S} tokens does not match kv range [{kv_start}, {kv_end})")
        gs = grid_sizes[0].tolist() if isinstance(grid_sizes, torch.Tensor) else list(grid_sizes[0])
        per_frame = gs[1] * gs[2]
        if kv_start % per_frame != 0:
            raise lib.B200Error("causvid: kv_start must be a whole number of latent frames")
        kc = self.kv_cache[block_idx]["k"].view(-1, D)
        vc = self.kv_cache[block_idx]["v"].view(-1, D)
        fp8 = self._is_fp8(weights.self_attn_q)
        n1 = self._ln(x, fp8, "a", scale=scale_msa, shift=shift_msa, eps=weights.norm1.eps)
        q = self._buf("cvq", (S, D), x.device)
        self._linear(weights.self_attn_q, n1, out=q)
        self._linear(weights.self_attn_k, n1, out=kc[kv_start:kv_end])
        self._linear(weights.self_attn_v, n1, out=vc[kv_start:kv_end])
        # RoPE rows of frames [start_frame, start_frame + f): table of the grid (start_frame + f, h, w), indexed at row + kv_start
        cs = self._rope_table([[kv_start // per_frame + gs[0], gs[1], gs[2]]], freqs, kv_end, x.device)
        lib.rms_rope_(q, weights.self_attn_norm_q.weight, kc[kv_start:kv_end], weights.self_attn_norm_k.weight, eps=weights.self_attn_norm_q.eps,
                      cos_sin=cs, rope_rows=S, pos_offset=kv_start)
        attn = self._buf("a", (S, D), x.device).view(S, H, d)
        lib.fmha(q.view(S, H, d), self.kv_cache[block_idx]["k"][:kv_end], self.kv_cache[block_idx]["v"][:kv_end], out=attn)
        self._linear(weights.self_attn_o, attn.reshape(S, D), out=x, epilogue=lib.EPI_GATE_RESIDUAL, gate=gate_msa)
        return x
