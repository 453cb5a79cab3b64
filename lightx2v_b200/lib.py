"""ctypes binding of ``libb200dit.so`` (C ABI declared in ``include/b200_dit.h``).

There is no fallback: if the library is missing or a call returns non-zero the caller gets an exception.
All wrappers take torch CUDA tensors, pass raw device pointers + the *current* torch stream, allocate nothing
inside the native code and never synchronise.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libb200dit.so")

# symbol -> (restype, argtypes); mirrors include/b200_dit.h one to one (tests check the header against this table)
_i64, _i32, _f32, _ptr = ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p
SIGNATURES = {
    "b200_last_error": (ctypes.c_char_p, []),
    "b200_version": (_i32, []),
    "b200_num_sms": (_i32, []),
    "b200_launch_count": (_i64, []),
    "b200_set_option": (_i32, [ctypes.c_char_p, _i32]),
    "b200_get_option": (_i32, [ctypes.c_char_p]),
    "b200_prof_fmha_begin": (_i32, [_i32]),
    "b200_prof_fmha_end": (_i32, [_ptr, _ptr, _i32]),
    "b200_fmha_fwd_d64": (_i32, [_ptr, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _i64, _i32, _f32, _ptr]),
    "b200_gemm_bf16": (_i32, [_ptr, _i64, _ptr, _i64, _ptr, _i64, _ptr, _ptr, _i64, _i64, _i64, _i32, _i32, _i32, _ptr]),
    "b200_ln_modulate": (_i32, [_ptr, _i64, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _f32, _ptr]),
    "b200_rms_rope": (_i32, [_ptr, _i64, _ptr, _ptr, _i64, _ptr, _i64, _i32, _f32, _ptr, _i64, _i64, _ptr]),
    "b200_fmha_fwd_d128": (_i32, [_ptr, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _i64, _i32, _f32, _ptr]),
    "b200_gemm_fp8": (_i32, [_ptr, _i64, _ptr, _i64, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i32, _i32, _i32, _ptr]),
    "b200_quant_fp8_per_token": (_i32, [_ptr, _i64, _ptr, _i64, _ptr, _i64, _i32, _ptr]),
    "b200_ln_modulate_fp8": (_i32, [_ptr, _i64, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _f32, _ptr]),
    "b200_rms_rope_heads": (_i32, [_ptr, _i64, _ptr, _ptr, _i64, _ptr, _i64, _i32, _f32, _ptr, _i64, _ptr]),
    "b200_ln_rope_heads64": (_i32, [_ptr, _i64, _ptr, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _i32, _f32, _ptr, _i64, _ptr]),
    "b200_debug_umma_rowshift": (_i32, [_ptr, _ptr, _ptr, _i32, _i32, _i32, _ptr]),
    "b200_debug_umma_rate": (_i32, [_i32, _i32, _i32, _i32, _i32, _i32, _i32, _ptr, _ptr]),
    "b200_rms_rope_scatter": (_i32, [_ptr, _i64, _ptr, _ptr, _i64, _i32, _f32, _ptr, _i64, _ptr, _i32, _i32, _i64, _ptr]),
    "b200_fmha_fwd_d128_scatter": (_i32, [_ptr, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i32, _i64, _i64, _i32, _i64, _i64, _i32, _f32, _ptr]),
    "b200_conv3d_cl": (_i32, [_ptr, _i64, _i64, _i64, _ptr, _ptr, _ptr, _i64, _i64, _i64, _ptr, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _i32,
                              _i32, _ptr, _i32, _ptr]),
    "b200_conv3d_cl_padded": (_i32, [_ptr, _i64, _i64, _i64, _i32, _i32, _i32, _ptr, _ptr, _ptr, _i64, _i64, _i64, _ptr, _i64, _i64, _i64, _i32, _i32,
                                     _i32, _i32, _i32, _i32, _ptr, _i32, _ptr]),
    "b200_wan_block_workspace_bytes": (_i64, [_i64, _i32, _i32]),
    "b200_wan_block_fwd": (_i32, [_ptr, _ptr, _ptr, _i64, _ptr]),
    "b200_quant_nvfp4": (_i32, [_ptr, _i64, _i64, _i32, _ptr, _ptr, _i64, _ptr, _ptr]),
    "b200_nvfp4_act_scale": (_i32, [_ptr, _i64, _i64, _i32, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "b200_gemm_nvfp4": (_i32, [_ptr, _i64, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _i64, _i64, _i32, _i32, _i32, _ptr]),
    "b200_gn_stats_workspace_doubles": (_i64, []),
    "b200_gn_stats_cl": (_i32, [_ptr, _i64, _i32, _ptr, _ptr]),
    "b200_gn_apply_pad_cl": (_i32, [_ptr, _ptr, _ptr, _ptr, _ptr, _f32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _ptr]),
    "b200_rms_silu_cl": (_i32, [_ptr, _ptr, _ptr, _i64, _i32, _i32, _ptr]),
    "b200_latent_to_cl": (_i32, [_ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _ptr]),
    "b200_cl_to_video": (_i32, [_ptr, _ptr, _i64, _i32, _i64, _ptr]),
}

EPI_BIAS, EPI_BIAS_GELU, EPI_GATE_RESIDUAL, EPI_RESIDUAL = 0, 1, 2, 3

_lib: Optional[ctypes.CDLL] = None


class B200Error(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the shared library (once). Raises if it has not been built — there is no CPU/torch fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). lightx2v_b200 has no fallback path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().b200_last_error()
        raise B200Error(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, name: str, dtype=torch.bfloat16) -> None:
    if not t.is_cuda:
        raise B200Error(f"{name}: expected a CUDA tensor (lightx2v_b200 has no CPU path)")
    if t.dtype != dtype:
        raise B200Error(f"{name}: expected {dtype}, got {t.dtype}")
    if t.stride(-1) != 1:
        raise B200Error(f"{name}: innermost dimension must be contiguous")


def gemm_bf16(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, out: Optional[torch.Tensor] = None,
              epilogue: int = EPI_BIAS, gate: Optional[torch.Tensor] = None, block_n: int = 0, max_ctas: int = 0) -> torch.Tensor:
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T + bias).  For the residual epilogues `out` is the in/out residual stream."""
    _req(a, "a"); _req(w, "w")
    M, K = a.shape
    N, K2 = w.shape
    if K != K2:
        raise B200Error(f"gemm_bf16: K mismatch {K} vs {K2}")
    if out is None:
        if epilogue in (EPI_GATE_RESIDUAL, EPI_RESIDUAL):
            raise B200Error("gemm_bf16: residual epilogues need out= (the residual stream, updated in place)")
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    _req(out, "out")
    rc = load().b200_gemm_bf16(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0),
                               _p(bias), _p(gate), M, N, K, epilogue, block_n, max_ctas, _stream())
    _check(rc, "b200_gemm_bf16")
    return out


def ln_modulate(x: torch.Tensor, *, weight=None, bias=None, scale=None, shift=None, eps: float = 1e-6,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(x, "x")
    rows, D = x.shape
    if out is None:
        out = torch.empty((rows, D), dtype=torch.bfloat16, device=x.device)
    rc = load().b200_ln_modulate(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), _p(weight), _p(bias),
                                 _p(scale), _p(shift), rows, D, eps, _stream())
    _check(rc, "b200_ln_modulate")
    return out


def rms_rope_(x0: torch.Tensor, w0: torch.Tensor, x1: Optional[torch.Tensor] = None, w1: Optional[torch.Tensor] = None, *,
              eps: float = 1e-6, cos_sin: Optional[torch.Tensor] = None, rope_rows: Optional[int] = None,
              pos_offset: int = 0) -> None:
    """In-place full-width RMSNorm (+ optional RoPE) on x0 (and x1)."""
    _req(x0, "x0")
    rows, D = x0.shape
    if x1 is not None:
        _req(x1, "x1")
        assert x1.shape == x0.shape
    if cos_sin is not None:
        _req(cos_sin, "cos_sin", torch.float32)
        assert cos_sin.is_contiguous() and cos_sin.shape[-2:] == (64, 2)
    rr = rows if rope_rows is None else rope_rows
    rc = load().b200_rms_rope(x0.data_ptr(), x0.stride(0), w0.data_ptr(), _p(x1), 0 if x1 is None else x1.stride(0),
                              _p(w1), rows, D, eps, _p(cos_sin), rr, pos_offset, _stream())
    _check(rc, "b200_rms_rope")


def fmha(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, softmax_scale: Optional[float] = None,
         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [Sq,H,d], k/v [Sk,H,d] (row-strided views allowed), d = 128 or 64 -> out [Sq,H,d]; one segment, non-causal."""
    d = q.shape[-1]
    if d not in (64, 128):
        raise B200Error(f"fmha: head_dim must be 64 or 128, got {d}")
    for n, t in (("q", q), ("k", k), ("v", v)):
        _req(t, n)
        if t.dim() != 3 or t.shape[2] != d or t.stride(1) != d:
            raise B200Error(f"fmha: {n} must be [S,H,{d}] with contiguous heads, got {tuple(t.shape)} / {t.stride()}")
    sq, H, _ = q.shape
    sk = k.shape[0]
    if out is None:
        out = torch.empty((sq, H, d), dtype=torch.bfloat16, device=q.device)
    _req(out, "out")
    scale = d ** -0.5 if softmax_scale is None else softmax_scale
    fn = load().b200_fmha_fwd_d128 if d == 128 else load().b200_fmha_fwd_d64
    rc = fn(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), out.data_ptr(), out.stride(0), sq, sk, H, scale, _stream())
    _check(rc, f"b200_fmha_fwd_d{d}")
    return out


def set_option(name: str, value: int) -> None:
    """Runtime A/B switch of the library ("conv_halo", "halo_base_offset", "conv_narrow")."""
    _check(load().b200_set_option(name.encode(), int(value)), "b200_set_option")


def get_option(name: str) -> int:
    return int(load().b200_get_option(name.encode()))


def launch_count() -> int:
    """Kernel launches issued by libb200dit.so in this process so far (bench.py's `gpu_launches` = difference over the timed region)."""
    return int(load().b200_launch_count())


def prof_fmha_begin(capacity: int = 8192) -> None:
    _check(load().b200_prof_fmha_begin(int(capacity)), "b200_prof_fmha_begin")


def prof_fmha_end(capacity: int = 8192):
    """-> list of (ms, sq, sk, heads, head_dim) for the attention launches since prof_fmha_begin (blocks until they finished)."""
    ms = (ctypes.c_float * capacity)()
    meta = (ctypes.c_int64 * (4 * capacity))()
    n = load().b200_prof_fmha_end(ctypes.cast(ms, ctypes.c_void_p), ctypes.cast(meta, ctypes.c_void_p), capacity)
    return [(float(ms[i]), int(meta[4 * i]), int(meta[4 * i + 1]), int(meta[4 * i + 2]), int(meta[4 * i + 3])) for i in range(n)]


# ---------------------------------------------------------------------------------------------------------------- fp8 (w8a8)
FP8 = torch.float8_e4m3fn


def quant_fp8_per_token(x: torch.Tensor, out: Optional[torch.Tensor] = None, scale: Optional[torch.Tensor] = None):
    """bf16 [rows, D] -> (e4m3 [rows, D], fp32 scale [rows, 1]), dynamic per-token (mm_weight.py:236-238)."""
    _req(x, "x")
    rows, D = x.shape
    if out is None:
        out = torch.empty((rows, D), dtype=FP8, device=x.device)
    if scale is None:
        scale = torch.empty((rows, 1), dtype=torch.float32, device=x.device)
    rc = load().b200_quant_fp8_per_token(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), scale.data_ptr(), rows, D, _stream())
    _check(rc, "b200_quant_fp8_per_token")
    return out, scale


def ln_modulate_fp8(x: torch.Tensor, *, weight=None, bias=None, scale=None, shift=None, eps: float = 1e-6,
                    out: Optional[torch.Tensor] = None, out_scale: Optional[torch.Tensor] = None):
    _req(x, "x")
    rows, D = x.shape
    if out is None:
        out = torch.empty((rows, D), dtype=FP8, device=x.device)
    if out_scale is None:
        out_scale = torch.empty((rows, 1), dtype=torch.float32, device=x.device)
    rc = load().b200_ln_modulate_fp8(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), out_scale.data_ptr(), _p(weight), _p(bias),
                                     _p(scale), _p(shift), rows, D, eps, _stream())
    _check(rc, "b200_ln_modulate_fp8")
    return out, out_scale


def gemm_fp8(a_q: torch.Tensor, a_scale: torch.Tensor, w_q: torch.Tensor, w_scale: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
             out: Optional[torch.Tensor] = None, epilogue: int = EPI_BIAS, gate: Optional[torch.Tensor] = None, block_n: int = 0,
             max_ctas: int = 0) -> torch.Tensor:
    """out[M,N] bf16 = epilogue(a_scale * (w_scale * (a_q[M,K] @ w_q[N,K]^T)) + bias); a_q, w_q e4m3."""
    _req(a_q, "a_q", FP8); _req(w_q, "w_q", FP8)
    _req(a_scale, "a_scale", torch.float32); _req(w_scale, "w_scale", torch.float32)
    M, K = a_q.shape
    N, K2 = w_q.shape
    if K != K2 or a_scale.numel() != M or w_scale.numel() != N:
        raise B200Error(f"gemm_fp8: shape mismatch a{tuple(a_q.shape)} w{tuple(w_q.shape)} sa{tuple(a_scale.shape)} sw{tuple(w_scale.shape)}")
    if out is None:
        if epilogue in (EPI_GATE_RESIDUAL, EPI_RESIDUAL):
            raise B200Error("gemm_fp8: residual epilogues need out= (the residual stream, updated in place)")
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a_q.device)
    _req(out, "out")
    rc = load().b200_gemm_fp8(a_q.data_ptr(), a_q.stride(0), w_q.data_ptr(), w_q.stride(0), out.data_ptr(), out.stride(0),
                              a_scale.data_ptr(), w_scale.data_ptr(), _p(bias), _p(gate), M, N, K, epilogue, block_n, max_ctas, _stream())
    _check(rc, "b200_gemm_fp8")
    return out


# ---------------------------------------------------------------------------------------------------------------- VAE (channels-last)
def conv3d_cl(x: torch.Tensor, wt: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor, taps, *, residual: Optional[torch.Tensor] = None,
              clamp_out: bool = False) -> torch.Tensor:
    """x, out, residual: channels-last VIEWS [T, H, W, C] (channels contiguous, arbitrary t/h/w strides, same T/H/W extent);
    wt [cout, ntaps*cin]; taps: sequence of (dt, dh, dw) input offsets."""
    _req(x, "x"); _req(wt, "wt"); _req(out, "out")
    T, H, W, cin = x.shape
    cout = wt.shape[0]
    if tuple(out.shape) != (T, H, W, cout) or wt.shape[1] != len(taps) * cin:
        raise B200Error(f"conv3d_cl: shape mismatch x{tuple(x.shape)} wt{tuple(wt.shape)} out{tuple(out.shape)} taps {len(taps)}")
    tp = (ctypes.c_int32 * (3 * len(taps)))(*[int(v) for t3 in taps for v in t3])
    rs = (0, 0, 0) if residual is None else (residual.stride(0), residual.stride(1), residual.stride(2))
    rc = load().b200_conv3d_cl(x.data_ptr(), x.stride(0), x.stride(1), x.stride(2), wt.data_ptr(), _p(bias), out.data_ptr(), out.stride(0),
                               out.stride(1), out.stride(2), _p(residual), rs[0], rs[1], rs[2], T, H, W, cin, cout, len(taps),
                               ctypes.cast(tp, ctypes.c_void_p), 1 if clamp_out else 0, _stream())
    _check(rc, "b200_conv3d_cl")
    return out


class WanBlockWeightsC(ctypes.Structure):
    """struct b200_wan_block_weights (include/b200_dit.h)."""
    _fields_ = [(n, ctypes.c_void_p) for n in ("wqkv", "bqkv", "norm_q", "norm_k", "wo", "bo", "norm3_w", "norm3_b", "wcq", "bcq", "cnorm_q", "wco", "bco",
                                               "w0", "b0", "w2", "b2")]


class WanBlockArgsC(ctypes.Structure):
    """struct b200_wan_block_args (include/b200_dit.h)."""
    _fields_ = ([("x", ctypes.c_void_p)] + [(n, ctypes.c_void_p) for n in ("shift_msa", "scale_msa", "gate_msa", "c_shift_msa", "c_scale_msa", "c_gate_msa")]
                + [("cos_sin", ctypes.c_void_p), ("rope_rows", ctypes.c_int64), ("ctx_k", ctypes.c_void_p), ("ctx_v", ctypes.c_void_p), ("ctx_len", ctypes.c_int64),
                   ("img_k", ctypes.c_void_p), ("img_v", ctypes.c_void_p), ("img_len", ctypes.c_int64), ("S", ctypes.c_int64), ("D", ctypes.c_int), ("H", ctypes.c_int),
                   ("F", ctypes.c_int), ("eps", ctypes.c_float)])


def wan_block_weights(**tensors) -> WanBlockWeightsC:
    """Pack bf16 CUDA tensors (kept alive by the caller) into the C struct."""
    w = WanBlockWeightsC()
    for name, _ in WanBlockWeightsC._fields_:
        t = tensors[name]
        _req(t, name)
        if not t.is_contiguous():
            raise B200Error(f"wan_block_weights: {name} must be contiguous")
        setattr(w, name, t.data_ptr())
    return w


def wan_block_workspace_bytes(S: int, D: int, F: int) -> int:
    return int(load().b200_wan_block_workspace_bytes(S, D, F))


def wan_block_fwd(w: WanBlockWeightsC, x: torch.Tensor, mods, cos_sin: torch.Tensor, rope_rows: int, ctx_k: torch.Tensor, ctx_v: torch.Tensor,
                  workspace: torch.Tensor, heads: int, ffn_dim: int, *, img_k: Optional[torch.Tensor] = None, img_v: Optional[torch.Tensor] = None,
                  eps: float = 1e-6) -> torch.Tensor:
    """x [S, D] <- Wan block(x) in one native call; mods = (shift_msa, scale_msa, gate_msa, c_shift_msa, c_scale_msa, c_gate_msa)."""
    _req(x, "x"); _req(ctx_k, "ctx_k"); _req(ctx_v, "ctx_v"); _req(cos_sin, "cos_sin", torch.float32)
    if not (x.is_contiguous() and ctx_k.is_contiguous() and ctx_v.is_contiguous()):
        raise B200Error("wan_block_fwd: x / ctx_k / ctx_v must be contiguous")
    S, D = x.shape
    a = WanBlockArgsC()
    a.x = x.data_ptr()
    for name, t in zip(("shift_msa", "scale_msa", "gate_msa", "c_shift_msa", "c_scale_msa", "c_gate_msa"), mods):
        _req(t, name)
        if not t.is_contiguous() or t.numel() != D:
            raise B200Error(f"wan_block_fwd: {name} must be a contiguous [D] vector")
        setattr(a, name, t.data_ptr())
    a.cos_sin, a.rope_rows = cos_sin.data_ptr(), rope_rows
    a.ctx_k, a.ctx_v, a.ctx_len = ctx_k.data_ptr(), ctx_v.data_ptr(), ctx_k.shape[0]
    if img_k is not None:
        _req(img_k, "img_k"); _req(img_v, "img_v")
        a.img_k, a.img_v, a.img_len = img_k.data_ptr(), img_v.data_ptr(), img_k.shape[0]
    else:
        a.img_k, a.img_v, a.img_len = None, None, 0
    a.S, a.D, a.H, a.F, a.eps = S, D, heads, ffn_dim, eps
    rc = load().b200_wan_block_fwd(ctypes.byref(w), ctypes.byref(a), workspace.data_ptr(), workspace.numel() * workspace.element_size(), _stream())
    _check(rc, "b200_wan_block_fwd")
    return x


def quant_nvfp4(x: torch.Tensor, global_scale: torch.Tensor):
    """bf16 [rows, K] -> (packed e2m1 uint8 [rows, K/2], ue4m3 scale factors uint8 [roundup(rows,128), K/16] in the 128x4 MMA layout)."""
    _req(x, "x"); _req(global_scale, "global_scale", torch.float32)
    rows, K = x.shape
    q = torch.empty((rows, K // 2), dtype=torch.uint8, device=x.device)
    sf = torch.empty(((rows + 127) // 128 * 128, K // 16), dtype=torch.uint8, device=x.device)
    rc = load().b200_quant_nvfp4(x.data_ptr(), x.stride(0), rows, K, global_scale.data_ptr(), q.data_ptr(), q.stride(0), sf.data_ptr(), _stream())
    _check(rc, "b200_quant_nvfp4")
    return q, sf


def nvfp4_act_scale(x: torch.Tensor, weight_global_scale: Optional[torch.Tensor] = None):
    """-> (global_scale fp32 [1] = 2688 / max|x|, alpha fp32 [1] = 1 / (global_scale * weight_global_scale)), computed on the device."""
    _req(x, "x")
    rows, K = x.shape
    buf = torch.empty(3, dtype=torch.float32, device=x.device)          # [global_scale, alpha, scratch]
    rc = load().b200_nvfp4_act_scale(x.data_ptr(), x.stride(0), rows, K, _p(weight_global_scale), buf[0:1].data_ptr(), buf[1:2].data_ptr(),
                                     buf[2:3].data_ptr(), _stream())
    _check(rc, "b200_nvfp4_act_scale")
    return buf[0:1], buf[1:2]


def gemm_nvfp4(a_q: torch.Tensor, b_q: torch.Tensor, sfa: torch.Tensor, sfb: torch.Tensor, alpha: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
               out: Optional[torch.Tensor] = None, epilogue: int = EPI_BIAS, gate: Optional[torch.Tensor] = None, block_n: int = 0,
               max_ctas: int = 0) -> torch.Tensor:
    """out[M,N] = epilogue(alpha * dequant(a_q, sfa) @ dequant(b_q, sfb)^T + bias)  (cutlass_scaled_fp4_mm semantics)."""
    _req(a_q, "a_q", torch.uint8); _req(b_q, "b_q", torch.uint8); _req(sfa, "sfa", torch.uint8); _req(sfb, "sfb", torch.uint8)
    _req(alpha, "alpha", torch.float32)
    M, K2 = a_q.shape
    N, K2b = b_q.shape
    if K2 != K2b:
        raise B200Error(f"gemm_nvfp4: K mismatch {K2 * 2} vs {K2b * 2}")
    if out is None:
        if epilogue in (EPI_GATE_RESIDUAL, EPI_RESIDUAL):
            raise B200Error("gemm_nvfp4: residual epilogues need out= (the residual stream, updated in place)")
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a_q.device)
    _req(out, "out")
    rc = load().b200_gemm_nvfp4(a_q.data_ptr(), a_q.stride(0), b_q.data_ptr(), b_q.stride(0), sfa.data_ptr(), sfb.data_ptr(), alpha.data_ptr(),
                                out.data_ptr(), out.stride(0), _p(bias), _p(gate), M, N, K2 * 2, epilogue, block_n, max_ctas, _stream())
    _check(rc, "b200_gemm_nvfp4")
    return out


def conv3d_cl_padded(xp: torch.Tensor, wt: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor, taps, *,
                     residual: Optional[torch.Tensor] = None, clamp_out: bool = False) -> torch.Tensor:
    """xp: pre-padded channels-last view [Tp, Hp, Wp, cin]; out: view [T, H, W, cout]; taps: non-negative (dt, dh, dw) offsets into xp."""
    _req(xp, "xp"); _req(wt, "wt"); _req(out, "out")
    Tp, Hp, Wp, cin = xp.shape
    T, H, W, cout = out.shape
    if wt.shape[0] != cout or wt.shape[1] != len(taps) * cin:
        raise B200Error(f"conv3d_cl_padded: shape mismatch xp{tuple(xp.shape)} wt{tuple(wt.shape)} out{tuple(out.shape)} taps {len(taps)}")
    tp = (ctypes.c_int32 * (3 * len(taps)))(*[int(v) for t3 in taps for v in t3])
    rs = (0, 0, 0) if residual is None else (residual.stride(0), residual.stride(1), residual.stride(2))
    rc = load().b200_conv3d_cl_padded(xp.data_ptr(), xp.stride(0), xp.stride(1), xp.stride(2), Tp, Hp, Wp, wt.data_ptr(), _p(bias), out.data_ptr(),
                                      out.stride(0), out.stride(1), out.stride(2), _p(residual), rs[0], rs[1], rs[2], T, H, W, cin, cout, len(taps),
                                      ctypes.cast(tp, ctypes.c_void_p), 1 if clamp_out else 0, _stream())
    _check(rc, "b200_conv3d_cl_padded")
    return out


def gn_stats_cl(x: torch.Tensor, sums: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-group (32 groups) sum / sum-of-squares of a contiguous channels-last tensor [..., C] -> fp64 workspace whose first 64
    entries are the result (the rest holds the per-block partials of the deterministic reduction)."""
    _req(x, "x")
    if not x.is_contiguous():
        raise B200Error("gn_stats_cl: x must be contiguous channels-last")
    C = x.shape[-1]
    need = int(load().b200_gn_stats_workspace_doubles())
    if sums is None:
        sums = torch.empty(need, dtype=torch.float64, device=x.device)
    elif sums.numel() < need:
        raise B200Error(f"gn_stats_cl: workspace too small ({sums.numel()} < {need} doubles)")
    rc = load().b200_gn_stats_cl(x.data_ptr(), x.numel() // C, C, sums.data_ptr(), _stream())
    _check(rc, "b200_gn_stats_cl")
    return sums


def gn_apply_pad_cl(x: torch.Tensor, sums: Optional[torch.Tensor], gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor], *, eps: float = 1e-6,
                    pad=(0, 0, 0), silu: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [T, H, W, C] -> [T + pt, H + 2 ph, W + 2 pw, C]: GroupNorm(32) affine [+ SiLU] with a replicate border (sums None: copy)."""
    _req(x, "x")
    if not x.is_contiguous():
        raise B200Error("gn_apply_pad_cl: x must be contiguous channels-last")
    T, H, W, C = x.shape
    pt, ph, pw = pad
    if out is None:
        out = torch.empty((T + pt, H + 2 * ph, W + 2 * pw, C), dtype=torch.bfloat16, device=x.device)
    if sums is not None:
        _req(gamma, "gamma", torch.float32); _req(beta, "beta", torch.float32)
    rc = load().b200_gn_apply_pad_cl(x.data_ptr(), out.data_ptr(), _p(sums), _p(gamma), _p(beta), float(eps), T, H, W, C, pt, ph, pw, 1 if silu else 0,
                                     _stream())
    _check(rc, "b200_gn_apply_pad_cl")
    return out


def rms_silu_cl(x: torch.Tensor, gamma: torch.Tensor, *, silu: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(x, "x"); _req(gamma, "gamma", torch.float32)
    if not x.is_contiguous():
        raise B200Error("rms_silu_cl: x must be contiguous channels-last")
    C = x.shape[-1]
    if out is None:
        out = torch.empty_like(x)
    rc = load().b200_rms_silu_cl(x.data_ptr(), out.data_ptr(), gamma.data_ptr(), x.numel() // C, C, 1 if silu else 0, _stream())
    _check(rc, "b200_rms_silu_cl")
    return out


def latent_to_cl(z: torch.Tensor, mean: torch.Tensor, inv_std: torch.Tensor, cp: int = 32) -> torch.Tensor:
    _req(z, "z", torch.float32)
    CZ, T, H, W = z.shape
    out = torch.empty((T, H, W, cp), dtype=torch.bfloat16, device=z.device)
    rc = load().b200_latent_to_cl(z.contiguous().data_ptr(), out.data_ptr(), mean.data_ptr(), inv_std.data_ptr(), T * H * W, CZ, cp, _stream())
    _check(rc, "b200_latent_to_cl")
    return out


def cl_to_video(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [T, H, W, CP] bf16 (3 valid channels) -> fp32 [3, T, H, W]; `out` may be a frame range `video[:, t0:t0 + T]` of a larger video."""
    _req(x, "x")
    T, H, W, CP = x.shape
    if out is None:
        out = torch.empty((3, T, H, W), dtype=torch.float32, device=x.device)
    if tuple(out.shape) != (3, T, H, W) or out.dtype != torch.float32 or out.stride(3) != 1 or out.stride(2) != W or out.stride(1) != H * W:
        raise B200Error(f"cl_to_video: out must be fp32 [3, {T}, {H}, {W}] with contiguous frames, got {tuple(out.shape)} / {out.stride()}")
    rc = load().b200_cl_to_video(x.contiguous().data_ptr(), out.data_ptr(), T * H * W, CP, out.stride(0), _stream())
    _check(rc, "b200_cl_to_video")
    return out


# ---------------------------------------------------------------------------------------------------------------- Ulysses over peer memory
def _ptr_table(ptrs):
    return (ctypes.c_void_p * len(ptrs))(*[int(p) for p in ptrs])


def rms_rope_scatter(qkv: torch.Tensor, wq: torch.Tensor, wk: torch.Tensor, cos_sin: torch.Tensor, peer_ptrs, rank: int, rows_per_rank: int,
                     *, eps: float = 1e-6, rope_rows: Optional[int] = None) -> None:
    """qkv [rows, 3*D] (local shard) -> RMSNorm+RoPE(q,k), copy(v), stored into each head owner's buffer peer_ptrs[r]
    laid out [world*rows_per_rank, 3, H/world, 128]."""
    _req(qkv, "qkv")
    rows, D3 = qkv.shape
    D = D3 // 3
    tbl = _ptr_table(peer_ptrs)
    rr = rows if rope_rows is None else rope_rows
    rc = load().b200_rms_rope_scatter(qkv.data_ptr(), qkv.stride(0), wq.data_ptr(), wk.data_ptr(), rows, D, eps, cos_sin.data_ptr(), rr,
                                      ctypes.cast(tbl, ctypes.c_void_p), len(peer_ptrs), rank, rows_per_rank, _stream())
    _check(rc, "b200_rms_rope_scatter")


def fmha_scatter(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, peer_ptrs, rows_per_rank: int, peer_stride_s: int, head_offset: int,
                 *, softmax_scale: Optional[float] = None) -> None:
    """Attention over all tokens for this rank's heads; output rows stored into the token owners' buffers."""
    for n, t in (("q", q), ("k", k), ("v", v)):
        _req(t, n)
    sq, H, _ = q.shape
    scale = 128 ** -0.5 if softmax_scale is None else softmax_scale
    tbl = _ptr_table(peer_ptrs)
    rc = load().b200_fmha_fwd_d128_scatter(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
                                           ctypes.cast(tbl, ctypes.c_void_p), len(peer_ptrs), rows_per_rank, peer_stride_s, head_offset,
                                           sq, k.shape[0], H, scale, _stream())
    _check(rc, "b200_fmha_fwd_d128_scatter")


# ---------------------------------------------------------------------------------------------------------------- HunyuanVideo q/k path
def rms_rope_heads_(x0: torch.Tensor, w0: torch.Tensor, x1: Optional[torch.Tensor] = None, w1: Optional[torch.Tensor] = None, *, eps: float = 1e-6,
                    cos_sin: Optional[torch.Tensor] = None, rope_rows: int = 0) -> None:
    """In place: per-head RMSNorm (weight [128]) of x0/x1 [rows, H, 128] (row-strided views), bf16-chain RoPE on rows < rope_rows."""
    _req(x0, "x0")
    rows, H, d = x0.shape
    if d != 128 or x0.stride(1) != 128:
        raise B200Error("rms_rope_heads_: expected [rows, H, 128] with contiguous heads")
    if cos_sin is not None:
        _req(cos_sin, "cos_sin", torch.float32)
    rc = load().b200_rms_rope_heads(x0.data_ptr(), x0.stride(0), w0.data_ptr(), _p(x1), 0 if x1 is None else x1.stride(0), _p(w1), rows, H, eps,
                                    _p(cos_sin), rope_rows if cos_sin is not None else 0, _stream())
    _check(rc, "b200_rms_rope_heads")


# ---------------------------------------------------------------------------------------------------------------- CogVideoX q/k path
def ln_rope_heads64_(x0: torch.Tensor, w0: torch.Tensor, b0: torch.Tensor, x1: Optional[torch.Tensor] = None, w1: Optional[torch.Tensor] = None,
                     b1: Optional[torch.Tensor] = None, *, eps: float = 1e-6, cos_sin: Optional[torch.Tensor] = None, rope_start: int = 0) -> None:
    """In place: per-head affine LayerNorm (weight / bias [64]) of x0/x1 [rows, H, 64] (row-strided views), pair rotation on rows >= rope_start
    with cos_sin [rows - rope_start, 32, 2] fp32."""
    _req(x0, "x0"); _req(w0, "w0"); _req(b0, "b0")
    rows, H, d = x0.shape
    if d != 64 or x0.stride(1) != 64:
        raise B200Error("ln_rope_heads64_: expected [rows, H, 64] with contiguous heads")
    if x1 is not None:
        _req(x1, "x1"); _req(w1, "w1"); _req(b1, "b1")
        if x1.shape != x0.shape or x1.stride(1) != 64:
            raise B200Error("ln_rope_heads64_: x1 must match x0's shape")
    if cos_sin is not None:
        _req(cos_sin, "cos_sin", torch.float32)
        if not cos_sin.is_contiguous() or tuple(cos_sin.shape[-2:]) != (32, 2) or cos_sin.shape[0] < rows - rope_start:
            raise B200Error(f"ln_rope_heads64_: cos_sin must be contiguous [>= {rows - rope_start}, 32, 2], got {tuple(cos_sin.shape)}")
    rc = load().b200_ln_rope_heads64(x0.data_ptr(), x0.stride(0), w0.data_ptr(), b0.data_ptr(), _p(x1), 0 if x1 is None else x1.stride(0), _p(w1), _p(b1),
                                     rows, H, eps, _p(cos_sin), rope_start, _stream())
    _check(rc, "b200_ln_rope_heads64")
