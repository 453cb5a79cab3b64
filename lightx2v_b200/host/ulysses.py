"""Ulysses sequence parallelism for the Wan block stack: token axis sharded across P ranks for every per-token op,
head-scatter / sequence-gather all-to-all around self-attention only.

Reference behaviour mirrored:
  pre_process / post_process  lightx2v/attentions/distributed/utils/wan/processor.py:9-37  (pad to a multiple of P, chunk, all_gather)
  ulysses_attn                lightx2v/attentions/distributed/ulysses/attn.py:7-91          (3 a2a in, attention on H/P heads, 1 a2a out)
  all2all_seq2head/head2seq   lightx2v/attentions/distributed/comm/all2all.py:7-89
  parallelize_wan             lightx2v/attentions/distributed/ulysses/wrap.py:53-71

B200 changes: q, k, v travel in ONE packed all-to-all ([P, S/P, 3, H/P, d], 290 MB per rank at 14B/720p/P=8)
instead of three; the received buffer is consumed in place by the FMHA as strided [S, H/P, d] views (no .contiguous()
transposes); no host synchronisation (the reference calls torch.cuda.synchronize() twice per block, attn.py:48,85);
padded key rows are masked out by passing the true sequence length to the FMHA (the reference lets zero-input pad
rows act as keys, SURVEY.md §5 "Long-context").
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def shard_rows(total: int, world: int) -> int:
    """Rows per rank after padding to a multiple of `world` (processor.py:13-19)."""
    return (total + world - 1) // world


def pre_process(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """processor.py:9-21: zero-pad the token axis to a multiple of P and keep this rank's contiguous chunk."""
    s = shard_rows(x.shape[0], world)
    pad = s * world - x.shape[0]
    if pad > 0:
        x = torch.cat([x, x.new_zeros(pad, x.shape[1])], dim=0)
    return x[rank * s:(rank + 1) * s].contiguous()


def post_process(x: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """processor.py:24-37: all_gather the shards along the token axis (padding rows dropped here rather than in unpatchify)."""
    world = dist.get_world_size(group)
    out = torch.empty((world * x.shape[0], x.shape[1]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out[:total]


def pack_qkv_for_a2a(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, world: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[s, H, d] x3 (any row stride) -> send buffer [P, s, 3, H/P, d]: block p holds heads [p*H/P, (p+1)*H/P) of q, k, v."""
    s, H, d = q.shape
    hp = H // world
    if out is None:
        out = torch.empty((world, s, 3, hp, d), dtype=q.dtype, device=q.device)
    for i, t in enumerate((q, k, v)):
        out[:, :, i].copy_(t.reshape(s, world, hp, d).permute(1, 0, 2, 3))
    return out


def unpack_out_from_a2a(recv: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """recv [P(src = head group), s, H/P, d] -> [s, H, d] with heads in global order (all2all.py:84-87)."""
    world, s, hp, d = recv.shape
    if out is None:
        out = torch.empty((s, world * hp, d), dtype=recv.dtype, device=recv.device)
    out.view(s, world, hp, d).copy_(recv.permute(1, 0, 2, 3))
    return out


class UlyssesAttention:
    """parallel_attention hook for WanTransformerInfer: q, k, v [S/P, H, d] -> out [S/P, H, d]."""

    def __init__(self, attention_fn: Callable, total_rows: int, group=None):
        self.attn = attention_fn            # (q[S,h,d], k, v, out=[S,h,d]) -> out   (lib.fmha on the GPU)
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.total_rows = total_rows        # true (unpadded) sequence length: padded keys are excluded
        self._bufs = {}

    def _buf(self, name, shape, like):
        key = (name, tuple(shape), like.dtype, str(like.device))
        b = self._bufs.get(key)
        if b is None:
            b = torch.empty(shape, dtype=like.dtype, device=like.device)
            self._bufs[key] = b
        return b

    def __call__(self, q, k, v, out=None, **_):
        P = self.world
        s, H, d = q.shape
        if H % P != 0:
            raise ValueError(f"Ulysses needs num_heads ({H}) divisible by world size ({P})")   # attn.py:37
        hp = H // P
        send = pack_qkv_for_a2a(q, k, v, P, self._buf("send", (P, s, 3, hp, d), q))
        recv = self._buf("recv", (P, s, 3, hp, d), q)
        dist.all_to_all_single(recv, send, group=self.group)
        full = recv.view(P * s, 3, hp, d)[: self.total_rows]              # rank-major == token order; drop pad rows
        osend = self._buf("osend", (P, s, hp, d), q)
        oview = osend.view(P * s, hp, d)
        if self.total_rows < P * s:
            oview[self.total_rows:].zero_()
        # [S, H/P, d] for all tokens, this rank's heads, written straight into the send buffer of the return exchange
        self.attn(full[:, 0], full[:, 1], full[:, 2], out=oview[: self.total_rows])
        orecv = self._buf("orecv", (P, s, hp, d), q)
        dist.all_to_all_single(orecv, osend, group=self.group)
        return unpack_out_from_a2a(orecv, out)


class FusedUlyssesAttention:
    """Ulysses exchange fused INTO the kernels over NVLink peer memory (no NCCL on the data path, no pack/unpack copies):

      rms_rope_scatter : q/k RMSNorm + RoPE (+ v copy) of the local shard, each 16-byte vector stored straight into the
                         receive buffer of the rank that owns its head            (replaces 3 all-to-alls + transposes)
      fmha_scatter     : attention over all tokens for the local heads, the epilogue stores each output row straight into
                         the buffer of the rank that owns the token               (replaces the return all-to-all)

    Buffers are torch symmetric-memory allocations (peer-mapped through NVSwitch); two stream-ordered cross-GPU barriers per
    block order the exchanges (writes of everyone -> my reads; my reads -> everyone's next writes).  Interface: called by
    WanTransformerInfer with the fused QKV buffer instead of q/k/v."""

    def __init__(self, total_rows: int, rows_per_rank: int, num_heads: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        from .. import lib

        self.lib = lib
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if num_heads % self.world != 0:
            raise ValueError(f"Ulysses needs num_heads ({num_heads}) divisible by world size ({self.world})")
        self.total_rows, self.s, self.H = total_rows, rows_per_rank, num_heads
        self.hp = num_heads // self.world
        # receive buffer of the head scatter: all tokens (incl. padding rows), q|k|v of my heads
        self.recv = symm_mem.empty((self.world * self.s, 3, self.hp, 128), dtype=torch.bfloat16, device=device)
        # attention output for my tokens, all heads (written by every rank)
        self.attn = symm_mem.empty((self.s, num_heads, 128), dtype=torch.bfloat16, device=device)
        # rows >= total_rows (padding to a multiple of the world size) are never written by the kernels: keep them zero, like the NCCL path,
        # so that nothing downstream (o-proj residual epilogue, the nvfp4 per-tensor absmax) ever sees uninitialised memory
        self.recv.zero_()
        self.attn.zero_()
        self.recv_h = symm_mem.rendezvous(self.recv, self.group)
        self.attn_h = symm_mem.rendezvous(self.attn, self.group)
        self.recv_ptrs = list(self.recv_h.buffer_ptrs)
        self.attn_ptrs = list(self.attn_h.buffer_ptrs)
        self.fused = True

    def run(self, qkv, wq, wk, cos_sin, eps) -> torch.Tensor:
        """qkv [s, 3*D] local shard (pre-norm) -> attention output for my tokens [s, H, 128] (a view of the symmetric buffer)."""
        lib = self.lib
        lib.rms_rope_scatter(qkv, wq, wk, cos_sin, self.recv_ptrs, self.rank, self.s, eps=eps)
        self.recv_h.barrier(channel=0)                                   # everyone's q/k/v for my heads have landed
        full = self.recv[: self.total_rows]
        lib.fmha_scatter(full[:, 0], full[:, 1], full[:, 2], self.attn_ptrs, self.s, self.H * 128, self.rank * self.hp)
        self.attn_h.barrier(channel=1)                                   # everyone's outputs for my tokens have landed
        return self.attn


def parallelize_wan_fused(model, total_rows: int, group=None):
    """Install the peer-memory Ulysses path on a host.wan_model.WanModel."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ti = model.transformer_infer
    s = shard_rows(total_rows, world)
    ti.parallel_attention = FusedUlyssesAttention(total_rows, s, ti.num_heads, torch.device("cuda", torch.cuda.current_device()), group)
    ti.sp_rank, ti.sp_world = rank, world
    model.pre_process = lambda x: pre_process(x, rank, world)
    model.post_process = lambda x: post_process(x, total_rows, group)
    return model


def parallelize_wan(model, total_rows: int, attention_fn: Callable, group=None):
    """Install the SP hooks on a host.wan_model.WanModel (wrap.py:53-71 does this by monkey-patching)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ti = model.transformer_infer
    ti.parallel_attention = UlyssesAttention(attention_fn, total_rows, group)
    ti.sp_rank, ti.sp_world = rank, world
    model.pre_process = lambda x: pre_process(x, rank, world)
    model.post_process = lambda x: post_process(x, total_rows, group)
    return model


def parallelize_wan_cfg(model, total_rows: int, attention_fn: Optional[Callable] = None, sp: str = "fused", group=None):
    """CFG-parallel x Ulysses: ranks [0, P/2) run the conditional pass, ranks [P/2, P) the unconditional one (SURVEY.md 8f N1;
    the reference runs them back to back, wan/model.py:203-218); inside each half the token axis is Ulysses-sharded over P/2 ranks
    (P = 2: no exchange at all inside the block stack).  `sp` selects the exchange inside a half ("fused": peer-memory kernels,
    "nccl": all-to-all); a failing peer-memory setup raises - there is no silent change of path.  Returns a description of what was installed."""
    if attention_fn is None:
        from .. import lib
        attention_fn = lib.fmha
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world % 2 != 0:
        raise ValueError(f"CFG-parallel needs an even world size, got {world}")
    half = world // 2
    base = dist.get_process_group_ranks(group) if group is not None else list(range(world))
    groups = [dist.new_group([base[i] for i in range(b * half, (b + 1) * half)]) for b in (0, 1)]      # collective: every rank creates both
    branch = rank // half
    model.cfg_parallel = (branch == 0, 0, half, group)
    mode = "cfg2"
    if half > 1:
        sub = groups[branch]
        if sp == "fused":
            parallelize_wan_fused(model, total_rows, sub)
            mode = f"cfg2 x ulysses{half} (peer memory)"
        else:
            parallelize_wan(model, total_rows, attention_fn, sub)
            mode = f"cfg2 x ulysses{half} (nccl)"
    return mode


class HunyuanUlyssesAttention:
    """parallel_attention hook for HunyuanTransformerInfer (reference: ulysses_attn with img_qkv_len / cu_seqlens_qkv,
    lightx2v/attentions/distributed/ulysses/attn.py:7-91; installed by parallelize_hunyuan, ulysses/wrap.py:5-50).

    Image tokens are sharded along the token axis (s rows per rank), text tokens are replicated.  One packed all-to-all gives every
    rank ALL image tokens for its H/P heads (:44-48); text q/k/v are sliced to the same heads locally (:51-53); attention runs over
    [image ; text] with the two varlen segments [0, img + txt_valid) and [img + txt_valid, img + txt_len) (:60-70); image outputs go
    back through the inverse all-to-all (:84-86), text outputs are all-gathered along heads (:80-81, 89)."""

    def __init__(self, attention_fn: Callable, group=None):
        self.attn = attention_fn
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._bufs = {}

    _buf = UlyssesAttention._buf

    def __call__(self, qkv3: torch.Tensor, img_len: int, bounds, out: torch.Tensor) -> torch.Tensor:
        """qkv3 [s + Lt, 3, H, d] (any row stride); bounds = global cu_seqlens; out [s + Lt, H, d] view."""
        P, r = self.world, self.rank
        L, _, H, d = qkv3.shape
        s, Lt = img_len, L - img_len
        if H % P != 0:
            raise ValueError(f"Ulysses needs num_heads ({H}) divisible by world size ({P})")
        hp = H // P
        img_total = P * s
        if bounds[-1] != img_total + Lt:
            raise ValueError(f"cu_seqlens {bounds} do not describe {img_total} image + {Lt} text tokens")
        send = self._buf("h_send", (P, s, 3, hp, d), qkv3)
        send.copy_(qkv3[:s].reshape(s, 3, P, hp, d).permute(2, 0, 1, 3, 4))
        full = self._buf("h_full", (img_total + Lt, 3, hp, d), qkv3)               # [all image tokens ; text], this rank's heads
        dist.all_to_all_single(full[:img_total].view(P, s, 3, hp, d), send, group=self.group)
        full[img_total:].copy_(qkv3[s:, :, r * hp:(r + 1) * hp])
        oall = self._buf("h_oall", (img_total + Lt, hp, d), qkv3)
        for a, b in zip(bounds[:-1], bounds[1:]):
            if b > a:
                self.attn(full[a:b, 0], full[a:b, 1], full[a:b, 2], out=oall[a:b])
        orecv = self._buf("h_orecv", (P, s, hp, d), qkv3)
        dist.all_to_all_single(orecv, oall[:img_total].view(P, s, hp, d), group=self.group)
        out[:s].copy_(orecv.permute(1, 0, 2, 3).reshape(s, H, d))
        tg = self._buf("h_txt", (P * Lt, hp, d), qkv3)
        dist.all_gather_into_tensor(tg, oall[img_total:].contiguous(), group=self.group)
        out[s:].copy_(tg.view(P, Lt, hp, d).permute(1, 0, 2, 3).reshape(Lt, H, d))
        return out


def parallelize_hunyuan(transformer_infer, attention_fn: Callable, group=None):
    """Install the Ulysses hook on a HunyuanTransformerInfer (wrap.py:5-50).  The caller shards the image tokens and the RoPE rows
    (contiguous chunks of the token axis; the reference splits the latent along H or W before patchify, utils/hunyuan/processor.py:5-50,
    which is the same thing up to the token order inside the shard) and all-gathers the image output."""
    transformer_infer.parallel_attention = HunyuanUlyssesAttention(attention_fn, group)
    return transformer_infer

