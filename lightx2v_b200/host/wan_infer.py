"""WanTransformerInfer with the reference's method surface, running each block as a short chain of fused sm_100a kernels.

Mirrors lightx2v/models/networks/wan/infer/transformer_infer.py (class WanTransformerInfer):
  infer(weights, grid_sizes, embed, x, embed0, seq_lens, freqs, context, audio_dit_blocks=None)   :80-81
  infer_block / infer_modulation / infer_self_attn / infer_cross_attn / infer_ffn / post_process     :289-508
and accepts either this package's weights tree (host/wan_weights.py) or the reference's own WanTransformerWeights
(the infer class only reads `.weight`, `.bias`, `.tensor`, `.eps` of the op objects).

Fusions per block (~50 launches / ~14 passes over [S,D] in the reference):
  LN + AdaLN modulate        -> b200_ln_modulate                       (1 read + 1 write of [S,D])
  q,k,v linears              -> ONE [S,D]x[D,3D] tcgen05 GEMM on the concatenated weight
  q/k RMSNorm + 3-axis RoPE  -> b200_rms_rope, in place on the fused QKV buffer (fp32 cos/sin table, no fp64 round trip)
  attention                  -> b200_fmha_fwd_d128 reading q/k/v as strided views of the QKV buffer
  o-proj + gate*y + residual -> GEMM epilogue (in place on x);  cross o-proj + residual -> GEMM epilogue
  ffn_0 + GELU(tanh)         -> GEMM epilogue;  ffn_2 + gate*y + residual -> GEMM epilogue
Text K/V of the cross-attention are step-invariant; they are cached per (block, context tensor).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .. import lib


def rope_cos_sin(grid_sizes, freqs: torch.Tensor, head_dim: int, rows: Optional[int] = None, row_offset: int = 0) -> torch.Tensor:
    """[rows, 64, 2] fp32 (cos, sin) table for the (f, h, w) token grid — the content of compute_freqs
    (lightx2v/models/networks/wan/infer/utils.py:7-20), or of compute_freqs_dist (:86-104) when rows/row_offset select a
    rank's shard (rows past f*h*w get the identity rotation, like pad_freqs' ones)."""
    c = head_dim // 2
    fs = freqs.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    f, h, w = [int(v) for v in grid_sizes]
    fi = torch.cat(
        [
            fs[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1),
            fs[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
            fs[2][:w].view(1, 1, w, -1).expand(f, h, w, -1),
        ],
        dim=-1,
    ).reshape(f * h * w, c)
    if rows is not None:
        total = row_offset + rows
        if total > fi.shape[0]:
            fi = torch.cat([fi, torch.ones(total - fi.shape[0], c, dtype=fi.dtype, device=fi.device)], dim=0)
        fi = fi[row_offset:total]
    return torch.stack([fi.real, fi.imag], dim=-1).to(torch.float32).contiguous()


class _BlockCache:
    """Per-block derived tensors: concatenated QKV weight/bias, cached cross-attention text K/V."""

    __slots__ = ("wqkv", "bqkv", "sqkv", "kv", "native", "keep", "owner", "fingerprint")

    def __init__(self):
        self.owner = None
        self.fingerprint = None
        self.wqkv = None
        self.bqkv = None
        self.sqkv = None
        self.kv = {}              # id(context tensor) -> (context tensor, its _version, ck, cv, ck_img, cv_img); see _context_kv
        self.native = None        # lib.WanBlockWeightsC + the tensors it points to
        self.keep = None


class WanTransformerInfer:
    def __init__(self, config):
        self.config = config
        self.task = config["task"]
        self.attention_type = config.get("attention_type", "b200_fmha")
        self.blocks_num = config["num_layers"]
        self.phases_num = 4
        self.num_heads = config["num_heads"]
        self.head_dim = config["dim"] // config["num_heads"]
        if self.head_dim != 128:
            raise lib.B200Error(f"WanTransformerInfer(B200): head_dim must be 128, got {self.head_dim}")
        self.parallel_attention = None          # set by parallel.ulysses.parallelize_wan
        self.sp_rank, self.sp_world = 0, 1
        self.infer_conditional = True
        self.scheduler = None
        self.mask_map = None
        self.cache_cross_kv = bool(config.get("b200_cache_cross_kv", True))
        self.native_block = bool(config.get("b200_native_block", True))   # one C call per block (csrc/wan_block.cu) when eligible
        self._caches: Dict[int, _BlockCache] = {}
        self._rope: Dict[Tuple, torch.Tensor] = {}
        self._bufs: Dict[Tuple, torch.Tensor] = {}
        self.infer_func = self._infer_without_offload

    # ------------------------------------------------------------------ reference surface
    def set_scheduler(self, scheduler):
        """BaseTransformerInfer.set_scheduler (lightx2v/common/transformer_infer/transformer_infer.py; called by WanModel.set_scheduler,
        wan/model.py:185)."""
        self.scheduler = scheduler

    def clear_weight_caches(self):
        """Drop every tensor derived from the weights (concatenated QKV, the native block's pointer struct, cached text K/V): call after the
        weight tree was reloaded or moved (WeightModule.load / to_cuda), so no block keeps computing with - or pinning - the old tensors."""
        self._caches.clear()

    def switch_status(self):
        self.infer_conditional = not self.infer_conditional

    def infer(self, weights, grid_sizes, embed, x, embed0, seq_lens, freqs, context, audio_dit_blocks=None):
        return self.infer_func(weights, grid_sizes, embed, x, embed0, seq_lens, freqs, context, audio_dit_blocks)

    def _infer_without_offload(self, weights, grid_sizes, embed, x, embed0, seq_lens, freqs, context, audio_dit_blocks=None):
        for block_idx in range(self.blocks_num):
            x = self.infer_block(weights.blocks[block_idx], grid_sizes, embed, x, embed0, seq_lens, freqs, context)
        return x

    def infer_block(self, weights, grid_sizes, embed, x, embed0, seq_lens, freqs, context):
        mods = self.infer_modulation(weights.compute_phases[0], embed0)
        if self.native_block and self._native_eligible(weights):
            return self._infer_block_native(weights, grid_sizes, x, mods, freqs, context)
        shift_msa, scale_msa, gate_msa, c_shift_msa, c_scale_msa, c_gate_msa = mods
        x = self.infer_self_attn(weights.compute_phases[1], grid_sizes, x, seq_lens, freqs, shift_msa, scale_msa, gate_msa=gate_msa)
        x, attn_out = self.infer_cross_attn(weights.compute_phases[2], x, context, None, None)
        y = self.infer_ffn(weights.compute_phases[3], x, attn_out, c_shift_msa, c_scale_msa, c_gate_msa=c_gate_msa)
        return self.post_process(x, y, c_gate_msa)

    # ------------------------------------------------------------------ whole block below the C ABI (csrc/wan_block.cu)
    def _native_eligible(self, weights) -> bool:
        """bf16 linears, no sequence-parallel hook: the cases b200_wan_block_fwd covers (fp8 / nvfp4 / Ulysses compose the per-op entries)."""
        sa = weights.compute_phases[1]
        return self.parallel_attention is None and not self._is_fp8(sa.self_attn_q) and not self._is_f4(sa.self_attn_q)

    def _infer_block_native(self, weights, grid_sizes, x, mods, freqs, context):
        sa, ca, ff = weights.compute_phases[1], weights.compute_phases[2], weights.compute_phases[3]
        S, D = x.shape
        dev = x.device
        c = self._cache(sa)
        if c.wqkv is None:
            c.wqkv = torch.cat([self._nk(sa.self_attn_q), self._nk(sa.self_attn_k), self._nk(sa.self_attn_v)], dim=0).contiguous()
            c.bqkv = torch.cat([sa.self_attn_q.bias, sa.self_attn_k.bias, sa.self_attn_v.bias]).contiguous()
        if c.native is None:
            c.keep = dict(wqkv=c.wqkv, bqkv=c.bqkv, norm_q=sa.self_attn_norm_q.weight, norm_k=sa.self_attn_norm_k.weight, wo=self._nk(sa.self_attn_o),
                          bo=sa.self_attn_o.bias, norm3_w=ca.norm3.weight, norm3_b=ca.norm3.bias, wcq=self._nk(ca.cross_attn_q), bcq=ca.cross_attn_q.bias,
                          cnorm_q=ca.cross_attn_norm_q.weight, wco=self._nk(ca.cross_attn_o), bco=ca.cross_attn_o.bias, w0=self._nk(ff.ffn_0),
                          b0=ff.ffn_0.bias, w2=self._nk(ff.ffn_2), b2=ff.ffn_2.bias)
            c.native = lib.wan_block_weights(**c.keep)
        cc = self._cache(ca)
        ck, cv, ck_img, cv_img = self._context_kv(ca, context, cc)
        F_ = c.keep["w0"].shape[0]
        ws = self._buf("native_ws", (lib.wan_block_workspace_bytes(S, D, F_) // 2,), dev)
        cs = self._rope_table(grid_sizes, freqs, S, dev)
        img = dict(img_k=ck_img.reshape(-1, D), img_v=cv_img.reshape(-1, D)) if self.task == "i2v" else {}
        return lib.wan_block_fwd(c.native, x, tuple(m.contiguous() for m in mods), cs, min(S, cs.shape[0]), ck.reshape(-1, D), cv.reshape(-1, D), ws,
                                 self.num_heads, F_, eps=sa.self_attn_norm_q.eps, **img)

    def infer_modulation(self, weights, embed0):
        """transformer_infer.py:308-319 (embed0 [6, D]); returns six contiguous [D] vectors."""
        if embed0.dim() != 2:
            raise lib.B200Error("WanTransformerInfer(B200): per-token (diffusion-forcing) embed0 is out of scope")
        mod = (weights.modulation.tensor + embed0).reshape(6, -1)
        return tuple(mod[i] for i in range(6))

    # ------------------------------------------------------------------ helpers
    def _buf(self, name, shape, device):
        key = (name, tuple(shape), str(device))
        b = self._bufs.get(key)
        if b is None:
            b = torch.empty(shape, dtype=torch.bfloat16, device=device)
            self._bufs[key] = b
        return b

    @staticmethod
    def _weight_fingerprint(weights):
        """(data_ptr, _version) of every weight tensor of a phase: changes when a tensor is replaced, moved or edited in place."""
        fp = []
        for m in getattr(weights, "_modules", {}).values():
            for a in ("weight", "bias", "weight_scale", "tensor"):
                t = getattr(m, a, None)
                if isinstance(t, torch.Tensor):
                    fp.append((t.data_ptr(), t._version))
        return tuple(fp)

    def _cache(self, weights) -> _BlockCache:
        """Derived tensors of one phase object.  The entry holds a reference to the phase (so id() cannot be reused by another object while
        the entry lives) and a fingerprint of its weight tensors; a reload / LoRA merge / to_cpu-to_cuda round trip rebuilds the entry."""
        fp = self._weight_fingerprint(weights)
        c = self._caches.get(id(weights))
        if c is None or c.owner is not weights or c.fingerprint != fp:
            c = _BlockCache()
            c.owner, c.fingerprint = weights, fp
            self._caches[id(weights)] = c
        return c

    def _rope_table(self, grid_sizes, freqs, rows, device):
        gs = grid_sizes[0].tolist() if isinstance(grid_sizes, torch.Tensor) else list(grid_sizes[0])
        key = (tuple(gs), rows, self.sp_rank, self.sp_world, str(device))
        t = self._rope.get(key)
        if t is None:
            if self.sp_world > 1:
                t = rope_cos_sin(gs, freqs.cpu(), self.head_dim, rows=rows, row_offset=self.sp_rank * rows)
            else:
                t = rope_cos_sin(gs, freqs.cpu(), self.head_dim)
            t = t.to(device)
            self._rope[key] = t
        return t

    @staticmethod
    def _nk(mm) -> torch.Tensor:
        """[N,K] row-major storage behind an MM op (ours or the reference's `weight = ckpt.t()` view)."""
        w = mm.weight.t()
        return w if w.is_contiguous() else w.contiguous()

    # ------------------------------------------------------------------ bf16 / w8a8-fp8 dispatch
    @staticmethod
    def _is_fp8(mm) -> bool:
        return getattr(mm, "weight_scale", None) is not None and not getattr(mm, "is_nvfp4", False)

    @staticmethod
    def _is_f4(mm) -> bool:
        return getattr(mm, "is_nvfp4", False)

    def _ln(self, x, fp8, name, **kw):
        """LayerNorm(+modulate) into the bf16 scratch buffer, or (fp8 mode) straight into e4m3 + per-token scale."""
        S, D = x.shape
        if fp8:
            q = self._buf8(name, (S, D), x.device)
            sc = self._buff32(name + "_s", (S, 1), x.device)
            return lib.ln_modulate_fp8(x, out=q, out_scale=sc, **kw)
        return lib.ln_modulate(x, out=self._buf(name, (S, D), x.device), **kw)

    def _linear(self, mm, a, *, w=None, b=None, ws=None, out=None, epilogue=lib.EPI_BIAS, gate=None, qname="q8"):
        """y = epilogue(a @ W^T + b).  `a` is a bf16 tensor, or an (e4m3, scale) pair when the producer already quantised."""
        if mm is not None and self._is_f4(mm):                     # w4a4: the op owns its packed weight and scale factors
            return mm.apply(a, out=out, epilogue=epilogue, gate=gate)
        w = self._nk(mm) if w is None else w
        b = mm.bias if b is None and mm is not None else b
        if mm is not None and self._is_fp8(mm) or ws is not None:
            ws = mm.weight_scale if ws is None else ws
            if not isinstance(a, tuple):
                q = self._buf8(qname, tuple(a.shape), a.device)
                sc = self._buff32(qname + "_s", (a.shape[0], 1), a.device)
                a = lib.quant_fp8_per_token(a, out=q, scale=sc)
            return lib.gemm_fp8(a[0], a[1], w, ws, b, out=out, epilogue=epilogue, gate=gate)
        return lib.gemm_bf16(a, w, b, out=out, epilogue=epilogue, gate=gate)

    def _buf8(self, name, shape, device):
        key = (name, tuple(shape), str(device), "fp8")
        t = self._bufs.get(key)
        if t is None:
            t = torch.empty(shape, dtype=lib.FP8, device=device)
            self._bufs[key] = t
        return t

    def _buff32(self, name, shape, device):
        key = (name, tuple(shape), str(device), "f32")
        t = self._bufs.get(key)
        if t is None:
            t = torch.empty(shape, dtype=torch.float32, device=device)
            self._bufs[key] = t
        return t

    # ------------------------------------------------------------------ phases
    def infer_self_attn(self, weights, grid_sizes, x, seq_lens, freqs, shift_msa, scale_msa, gate_msa=None):
        """transformer_infer.py:321-396 (+ the gated residual of :402 when gate_msa is given: returns the updated x;
        with gate_msa=None returns y like the reference)."""
        S, D = x.shape
        dev = x.device
        c = self._cache(weights)
        fp8 = self._is_fp8(weights.self_attn_q)
        f4 = self._is_f4(weights.self_attn_q)
        if c.wqkv is None and not f4:
            c.wqkv = torch.cat([self._nk(weights.self_attn_q), self._nk(weights.self_attn_k), self._nk(weights.self_attn_v)], dim=0).contiguous()
            c.bqkv = torch.cat([weights.self_attn_q.bias, weights.self_attn_k.bias, weights.self_attn_v.bias]).contiguous()
            if fp8:
                c.sqkv = torch.cat([m.weight_scale.reshape(-1) for m in (weights.self_attn_q, weights.self_attn_k, weights.self_attn_v)]).contiguous()
        n1 = self._ln(x, fp8, "a", scale=scale_msa, shift=shift_msa, eps=weights.norm1.eps)
        qkv = self._buf("qkv", (S, 3 * D), dev)
        if f4:      # one activation quantisation shared by the three projections (each weight has its own global scale)
            xq3 = weights.self_attn_q.quantize_input(n1)
            for i, mm in enumerate((weights.self_attn_q, weights.self_attn_k, weights.self_attn_v)):
                mm.apply_q(xq3, out=qkv[:, i * D:(i + 1) * D])
        else:
            self._linear(None, n1, w=c.wqkv, b=c.bqkv, ws=c.sqkv, out=qkv)
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
        cs = self._rope_table(grid_sizes, freqs, S, dev)
        H, d = self.num_heads, self.head_dim
        if getattr(self.parallel_attention, "fused", False):
            # Ulysses over peer memory: norm + RoPE + head scatter in one kernel, attention with a token-scatter epilogue
            attn = self.parallel_attention.run(qkv, weights.self_attn_norm_q.weight, weights.self_attn_norm_k.weight, cs,
                                               weights.self_attn_norm_q.eps)
        else:
            lib.rms_rope_(q, weights.self_attn_norm_q.weight, k, weights.self_attn_norm_k.weight, eps=weights.self_attn_norm_q.eps,
                          cos_sin=cs, rope_rows=min(S, cs.shape[0]))
            q3, k3, v3 = (t.unflatten(1, (H, d)) for t in (q, k, v))
            attn = self._buf("a", (S, D), dev).view(S, H, d)   # the LN scratch is dead after the QKV GEMM: reuse it for the attention output
            if self.parallel_attention is None:
                lib.fmha(q3, k3, v3, out=attn)
            else:
                attn = self.parallel_attention(q=q3, k=k3, v=v3, out=attn)
        attn2 = attn.reshape(S, D)
        if gate_msa is None:
            return self._linear(weights.self_attn_o, attn2)
        self._linear(weights.self_attn_o, attn2, out=x, epilogue=lib.EPI_GATE_RESIDUAL, gate=gate_msa)
        return x

    def _context_kv(self, weights, context, c: _BlockCache):
        """Text (and CLIP) K/V of this block for `context`, computed once per context tensor (the reference recomputes them
        in every block of every step, transformer_infer.py:418-420).  The cache is keyed on the tensor OBJECT and keeps a reference to it, so
        a freed-and-reallocated buffer at the same address can never alias an entry; an in-place update bumps `_version` and invalidates it.
        Up to four contexts per block stay cached (cond / uncond alternate inside one step)."""
        ent = c.kv.get(id(context)) if self.cache_cross_kv else None
        if ent is not None and ent[0] is context and ent[1] == context._version:
            return ent[2:]
        H, d = self.num_heads, self.head_dim
        if self.task == "i2v":
            context_img, ctx = context[:257], context[257:]
        else:
            context_img, ctx = None, context
        ck = self._linear(weights.cross_attn_k, ctx.contiguous(), qname="ctx8")
        lib.rms_rope_(ck, weights.cross_attn_norm_k.weight, eps=weights.cross_attn_norm_k.eps)
        cv = self._linear(weights.cross_attn_v, ctx.contiguous(), qname="ctx8")
        out = [ck.view(-1, H, d), cv.view(-1, H, d), None, None]
        if context_img is not None:
            ki = self._linear(weights.cross_attn_k_img, context_img.contiguous(), qname="img8")
            lib.rms_rope_(ki, weights.cross_attn_norm_k_img.weight, eps=weights.cross_attn_norm_k_img.eps)
            vi = self._linear(weights.cross_attn_v_img, context_img.contiguous(), qname="img8")
            out[2], out[3] = ki.view(-1, H, d), vi.view(-1, H, d)
        if self.cache_cross_kv:
            if len(c.kv) >= 4:
                c.kv.pop(next(iter(c.kv)))
            c.kv[id(context)] = (context, context._version, *out)
        return tuple(out)

    def infer_cross_attn(self, weights, x, context, y_out, gate_msa):
        """transformer_infer.py:398-465.  If y_out is given the gated residual `x += y_out * gate_msa` (:402) is applied
        here like the reference; the fused path passes None because the self-attention o-proj epilogue already did it.
        Returns (x, attn_out) where attn_out is None when the cross o-proj epilogue already accumulated into x."""
        S, D = x.shape
        dev = x.device
        H, d = self.num_heads, self.head_dim
        if y_out is not None:
            x.add_(y_out * gate_msa)
        c = self._cache(weights)
        fp8 = self._is_fp8(weights.cross_attn_q)
        n3 = self._ln(x, fp8, "a", weight=weights.norm3.weight, bias=weights.norm3.bias, eps=weights.norm3.eps)
        cq = self._buf("b", (S, D), dev)
        self._linear(weights.cross_attn_q, n3, out=cq)
        lib.rms_rope_(cq, weights.cross_attn_norm_q.weight, eps=weights.cross_attn_norm_q.eps)
        ck, cv, ck_img, cv_img = self._context_kv(weights, context, c)
        attn = self._buf("a", (S, D), dev).view(S, H, d)
        lib.fmha(cq.view(S, H, d), ck, cv, out=attn)
        if self.task == "i2v":
            img = lib.fmha(cq.view(S, H, d), ck_img, cv_img, out=self._buf("c", (S, H, d), dev))
            attn.add_(img)                                                      # :454 (two softmaxes, summed in bf16)
        self._linear(weights.cross_attn_o, attn.reshape(S, D), out=x, epilogue=lib.EPI_RESIDUAL)
        return x, None

    def infer_ffn(self, weights, x, attn_out, c_shift_msa, c_scale_msa, c_gate_msa=None):
        """transformer_infer.py:467-497 (+ :503 when c_gate_msa is given: the ffn_2 epilogue updates x and None is returned)."""
        S, D = x.shape
        dev = x.device
        if attn_out is not None:
            x.add_(attn_out)
        fp8 = self._is_fp8(weights.ffn_0)
        n2 = self._ln(x, fp8, "a", scale=c_scale_msa, shift=c_shift_msa, eps=weights.norm2.eps)
        f4 = self._is_f4(weights.ffn_0)
        w0 = None if f4 else self._nk(weights.ffn_0)
        hidden = self._buf("h", (S, weights.ffn_0.out_features if f4 else w0.shape[0]), dev)
        self._linear(weights.ffn_0, n2, w=w0, out=hidden, epilogue=lib.EPI_BIAS_GELU)
        if c_gate_msa is None:
            return self._linear(weights.ffn_2, hidden, qname="h8")
        self._linear(weights.ffn_2, hidden, out=x, epilogue=lib.EPI_GATE_RESIDUAL, gate=c_gate_msa, qname="h8")
        return None

    def post_process(self, x, y, c_gate_msa):
        """transformer_infer.py:499-508."""
        if y is not None:
            x.add_(y * c_gate_msa)
        return x
