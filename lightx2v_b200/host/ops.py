"""Weight-holding op objects with LightX2V's operator interface (`load / apply / to_cuda / to_cpu / state_dict /
set_config / clear / _calculate_size`), backed by the sm_100a kernels of libb200dit.so.

Reference classes mirrored (same constructor arguments, same `apply` signatures, same state_dict keys):
  MMWeightB200    <- MMWeight         lightx2v/common/ops/mm/mm_weight.py:29-96
  RMSWeightB200   <- RMSWeightSgl     lightx2v/common/ops/norm/rms_norm_weight.py:12-118
  LNWeightB200    <- LNWeight         lightx2v/common/ops/norm/layer_norm_weight.py:7-111
  FmhaWeightB200  <- FlashAttn2Weight lightx2v/common/ops/attn/attn_weight.py:43-97
  DefaultTensor   <- DefaultTensor    lightx2v/common/ops/tensor/tensor.py:6-47
`apply` raises if the tensor is not on a CUDA device: there is no CPU / torch fallback on this path.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .. import lib
from .registry import ATTN_KEY, ATTN_WEIGHT_REGISTER, LN_WEIGHT_REGISTER, MM_KEY, MM_WEIGHT_REGISTER, RMS_WEIGHT_REGISTER, TENSOR_REGISTER


class _WeightOp:
    """to_cuda / to_cpu / clear plumbing shared by the weight holders (attribute names listed in `_attrs`)."""

    _attrs = ()

    def set_config(self, config=None):
        if config is not None:
            self.config = config

    def to_cuda(self, non_blocking=False):
        for a in self._attrs:
            t = getattr(self, a, None)
            if t is not None:
                setattr(self, a, t.cuda(non_blocking=non_blocking))

    def to_cpu(self, non_blocking=False):
        for a in self._attrs:
            t = getattr(self, a, None)
            if t is not None:
                setattr(self, a, t.to("cpu", non_blocking=non_blocking))

    def clear(self):
        for a in self._attrs:
            if hasattr(self, a):
                setattr(self, a, None)

    def _calculate_size(self):
        return sum(t.numel() * t.element_size() for t in (getattr(self, a, None) for a in self._attrs) if t is not None)


@MM_WEIGHT_REGISTER(MM_KEY)
class MMWeightB200(_WeightOp):
    """y = x @ W^T + b through the tcgen05 GEMM.  `self.weight` keeps the reference's convention — the transposed VIEW
    `[K, N]` of the checkpoint's `[N, K]` tensor (mm_weight.py:76) — so code that reads `.weight` sees the same thing;
    the kernel consumes the underlying `[N, K]` row-major storage (an NT GEMM)."""

    _attrs = ("weight", "bias")

    def __init__(self, weight_name, bias_name, lazy_load=False, lazy_load_file=None):
        self.weight_name = weight_name
        self.bias_name = bias_name
        self.lazy_load = lazy_load
        self.lazy_load_file = lazy_load_file
        self.config = {}
        self.weight = None
        self.bias = None

    def load(self, weight_dict: Dict[str, torch.Tensor]):
        w = weight_dict[self.weight_name]
        if w.dtype != torch.bfloat16:
            raise lib.B200Error(f"{self.weight_name}: B200-bf16 expects a bf16 checkpoint tensor, got {w.dtype}")
        self.weight = w.contiguous().t()
        self.bias = weight_dict[self.bias_name] if self.bias_name is not None else None

    def to_cuda(self, non_blocking=False):
        # .cuda() of a transposed view stays a transposed view of a contiguous [N,K] buffer
        self.weight = self.weight.t().cuda(non_blocking=non_blocking).t()
        if self.bias is not None:
            self.bias = self.bias.cuda(non_blocking=non_blocking)

    def to_cpu(self, non_blocking=False):
        self.weight = self.weight.t().to("cpu", non_blocking=non_blocking).t()
        if self.bias is not None:
            self.bias = self.bias.to("cpu", non_blocking=non_blocking)

    @property
    def weight_nk(self) -> torch.Tensor:
        return self.weight.t()

    def apply(self, input_tensor: torch.Tensor, *, out=None, epilogue: int = lib.EPI_BIAS, gate=None) -> torch.Tensor:
        return lib.gemm_bf16(input_tensor, self.weight.t(), self.bias, out=out, epilogue=epilogue, gate=gate)

    def state_dict(self, destination=None):
        if destination is None:
            destination = {}
        destination[self.weight_name] = self.weight.cpu().detach().clone().t().contiguous()
        if self.bias is not None:
            destination[self.bias_name] = self.bias.cpu().detach().clone()
        return destination


FP8_MM_KEY = "W-fp8-channel-sym-A-fp8-channel-sym-dynamic-B200"


def quantize_weight_fp8_per_channel(w: torch.Tensor):
    """Per-out-channel symmetric e4m3 weight quantisation — FloatQuantizer("e4m3", True, "per_channel").real_quant_tensor
    (lightx2v/utils/quant_utils.py:41-53,149-161; same rule as tools/convert/converter.py:313-339):
    scale[n] = clamp(absmax(w[n,:]), 1e-5) / 448;  q = round_to_nearest_e4m3(clip(w / scale, -448, 448))."""
    wf = w.to(torch.float32)
    absmax = wf.abs().amax(dim=-1, keepdim=True).clamp(min=1e-5)
    scale = absmax / 448.0
    q = torch.clip(wf / scale, -448.0, 448.0).to(torch.float8_e4m3fn)
    return q, scale.to(torch.float32)


@MM_WEIGHT_REGISTER(FP8_MM_KEY)
class MMWeightFp8B200(_WeightOp):
    """w8a8-fp8 linear (BASELINE config 3): e4m3 weight [N,K] + `weight_scale` [N,1] fp32 (the checkpoint contract of
    tools/convert/converter.py:294-339 and MMWeightQuantTemplate.load_quantized, mm_weight.py:161-166), dynamic per-token e4m3
    activations, `y = sx * (sw * (xq @ wq^T)) + b`.  Mirrors MMWeightWfp8channelAfp8channeldynamicVllm (mm_weight.py:287-319):
    `config["weight_auto_quant"]` quantises a bf16 checkpoint at load, otherwise `<name>.weight` must already be e4m3 and
    `<name>.weight_scale` present."""

    _attrs = ("weight", "weight_scale", "bias")

    def __init__(self, weight_name, bias_name, lazy_load=False, lazy_load_file=None):
        self.weight_name = weight_name
        self.bias_name = bias_name
        self.weight_scale_name = weight_name.removesuffix(".weight") + ".weight_scale"
        self.lazy_load = lazy_load
        self.lazy_load_file = lazy_load_file
        self.config = {}
        self.weight = self.weight_scale = self.bias = None

    def load(self, weight_dict):
        w = weight_dict[self.weight_name]
        if (self.config or {}).get("weight_auto_quant", False) or w.dtype != torch.float8_e4m3fn:
            if w.dtype == torch.float8_e4m3fn:
                raise lib.B200Error(f"{self.weight_name}: already e4m3 but weight_auto_quant requested")
            q, sc = quantize_weight_fp8_per_channel(w)
        else:
            q, sc = w, weight_dict[self.weight_scale_name].float()
        self.weight = q.contiguous().t()                       # [K,N] view, like the reference (weight_need_transpose)
        self.weight_scale = sc.reshape(-1, 1).contiguous()
        self.bias = weight_dict[self.bias_name] if self.bias_name is not None else None

    def to_cuda(self, non_blocking=False):
        self.weight = self.weight.t().cuda(non_blocking=non_blocking).t()
        self.weight_scale = self.weight_scale.cuda(non_blocking=non_blocking)
        if self.bias is not None:
            self.bias = self.bias.cuda(non_blocking=non_blocking)

    def to_cpu(self, non_blocking=False):
        self.weight = self.weight.t().to("cpu", non_blocking=non_blocking).t()
        self.weight_scale = self.weight_scale.to("cpu", non_blocking=non_blocking)
        if self.bias is not None:
            self.bias = self.bias.to("cpu", non_blocking=non_blocking)

    def apply(self, input_tensor, *, out=None, epilogue: int = lib.EPI_BIAS, gate=None):
        xq, sx = lib.quant_fp8_per_token(input_tensor)
        return lib.gemm_fp8(xq, sx, self.weight.t(), self.weight_scale, self.bias, out=out, epilogue=epilogue, gate=gate)

    def state_dict(self, destination=None):
        if destination is None:
            destination = {}
        destination[self.weight_name] = self.weight.cpu().detach().clone().t().contiguous()
        destination[self.weight_scale_name] = self.weight_scale.cpu().detach().clone()
        if self.bias is not None:
            destination[self.bias_name] = self.bias.cpu().detach().clone()
        return destination


NVFP4_MM_KEY = "W-nvfp4-group16-A-nvfp4-group16-dynamic-B200"


def quantize_weight_nvfp4(w: torch.Tensor):
    """bf16 [N, K] (CUDA) -> (packed e2m1 uint8 [N, K/2], ue4m3 scale bytes uint8 [roundup(N,128), K/16] in the MMA layout,
    global_scale fp32 [1] = 448 * 6 / max|w|) - the recipe of lightx2v_kernel/docs/en_US/nvfp4_quantization_basics.md:34-80 with the
    arithmetic of test/nvfp4_nvfp4/fake_quant.py.  Runs on the GPU (the quantiser is a CUDA kernel; there is no CPU path)."""
    w = w.to(torch.bfloat16).contiguous()
    gs, _ = lib.nvfp4_act_scale(w)
    q, sf = lib.quant_nvfp4(w, gs)
    return q, sf, gs.clone()


def quantize_checkpoint_nvfp4(weight_dict: Dict[str, torch.Tensor], names, device="cuda") -> Dict[str, torch.Tensor]:
    """Offline converter for the w4a4 path: for every `<name>.weight` in `names` emit
    `<name>.weight` (packed e2m1), `<name>.weight_scale` (swizzled ue4m3 bytes) and `<name>.weight_global_scale`;
    everything else is passed through unchanged."""
    out = dict(weight_dict)
    for n in names:
        q, sf, gs = quantize_weight_nvfp4(weight_dict[n].to(device))
        base = n.removesuffix(".weight")
        out[n], out[base + ".weight_scale"], out[base + ".weight_global_scale"] = q.cpu(), sf.cpu(), gs.cpu()
    return out


@MM_WEIGHT_REGISTER(NVFP4_MM_KEY)
class MMWeightNvfp4B200(_WeightOp):
    """w4a4 NVFP4 linear: packed e2m1 weight [N, K/2] + ue4m3 group scales + per-tensor global scale, dynamic per-tensor activation
    scale computed on the device, `y = alpha * (xq @ wq^T) + b` with alpha = 1 / (gs_x * gs_w)  (cutlass_scaled_fp4_mm semantics,
    lightx2v_kernel/python/lightx2v_kernel/gemm.py:4-52, test/nvfp4_nvfp4/test_bench1.py:106-138).  The reference has the kernels but
    no MM_WEIGHT class or checkpoint loader for nvfp4; this class follows the fp8 template's contract:
    a bf16 `<name>.weight` is quantised when it first reaches the GPU, a uint8 one must come with `weight_scale` and
    `weight_global_scale` (see quantize_checkpoint_nvfp4)."""

    is_nvfp4 = True
    _attrs = ("weight", "weight_scale", "weight_global_scale", "bias")

    def __init__(self, weight_name, bias_name, lazy_load=False, lazy_load_file=None):
        self.weight_name = weight_name
        self.bias_name = bias_name
        base = weight_name.removesuffix(".weight")
        self.weight_scale_name = base + ".weight_scale"
        self.weight_global_scale_name = base + ".weight_global_scale"
        self.lazy_load = lazy_load
        self.lazy_load_file = lazy_load_file
        self.config = {}
        self.weight = self.weight_scale = self.weight_global_scale = self.bias = None
        self._raw = None

    def load(self, weight_dict):
        w = weight_dict[self.weight_name]
        if w.dtype == torch.uint8:
            self.weight = w.contiguous()
            self.weight_scale = weight_dict[self.weight_scale_name].view(torch.uint8).contiguous()
            self.weight_global_scale = weight_dict[self.weight_global_scale_name].float().reshape(1)
        else:
            self._raw = w
            if w.is_cuda:
                self._quantize_raw()
        self.bias = weight_dict[self.bias_name] if self.bias_name is not None else None

    def _quantize_raw(self):
        self.weight, self.weight_scale, self.weight_global_scale = quantize_weight_nvfp4(self._raw.cuda())
        self._raw = None

    @property
    def out_features(self):
        return (self.weight if self.weight is not None else self._raw).shape[0]

    def to_cuda(self, non_blocking=False):
        if self._raw is not None:
            self._quantize_raw()
        super().to_cuda(non_blocking)

    def quantize_input(self, x: torch.Tensor):
        """-> (packed activations, scale factors, global_scale): reusable across several weights that share the input."""
        gs, _ = lib.nvfp4_act_scale(x)
        xq, sfx = lib.quant_nvfp4(x, gs)
        return xq, sfx, gs

    def apply_q(self, xq3, *, out=None, epilogue: int = lib.EPI_BIAS, gate=None, block_n: int = 0):
        xq, sfx, gs = xq3
        alpha = torch.reciprocal(gs * self.weight_global_scale)
        return lib.gemm_nvfp4(xq, self.weight, sfx, self.weight_scale, alpha, self.bias, out=out, epilogue=epilogue, gate=gate, block_n=block_n)

    def apply(self, input_tensor, *, out=None, epilogue: int = lib.EPI_BIAS, gate=None):
        if self.weight is None:
            raise lib.B200Error(f"{self.weight_name}: weight not on the GPU yet (call to_cuda(); nvfp4 has no CPU path)")
        gs, alpha = lib.nvfp4_act_scale(input_tensor, self.weight_global_scale)
        xq, sfx = lib.quant_nvfp4(input_tensor, gs)
        return lib.gemm_nvfp4(xq, self.weight, sfx, self.weight_scale, alpha, self.bias, out=out, epilogue=epilogue, gate=gate)

    def state_dict(self, destination=None):
        if destination is None:
            destination = {}
        destination[self.weight_name] = self.weight.cpu().detach().clone()
        destination[self.weight_scale_name] = self.weight_scale.cpu().detach().clone()
        destination[self.weight_global_scale_name] = self.weight_global_scale.cpu().detach().clone()
        if self.bias is not None:
            destination[self.bias_name] = self.bias.cpu().detach().clone()
        return destination


class RMSWeightB200(_WeightOp):
    """Full-row RMSNorm with the reference's bf16 rounding points (rms_norm_weight.py:111-113)."""

    _attrs = ("weight",)

    def __init__(self, weight_name, lazy_load=False, lazy_load_file=None, eps=1e-6):
        self.weight_name = weight_name
        self.eps = eps
        self.lazy_load = lazy_load
        self.lazy_load_file = lazy_load_file
        self.config = {}
        self.weight = None

    def load(self, weight_dict):
        if not self.lazy_load:
            self.weight = weight_dict[self.weight_name]

    def load_from_disk(self):
        self.weight = self.lazy_load_file.get_tensor(self.weight_name).to(torch.bfloat16)

    def apply(self, input_tensor: torch.Tensor) -> torch.Tensor:
        shape = input_tensor.shape
        x = input_tensor.reshape(-1, shape[-1]).clone()      # the reference returns a new tensor
        lib.rms_rope_(x, self.weight, eps=self.eps)
        return x.view(shape)

    def state_dict(self, destination=None):
        if destination is None:
            destination = {}
        destination[self.weight_name] = self.weight.cpu().detach().clone()
        return destination


RMS_WEIGHT_REGISTER["Default"] = RMSWeightB200
RMS_WEIGHT_REGISTER["sgl-kernel"] = RMSWeightB200


class LNWeightB200(_WeightOp):
    """LayerNorm, optional affine (layer_norm_weight.py:100-111)."""

    _attrs = ("weight", "bias")

    def __init__(self, weight_name=None, bias_name=None, lazy_load=False, lazy_load_file=None, eps=1e-6):
        self.weight_name = weight_name
        self.bias_name = bias_name
        self.eps = eps
        self.lazy_load = lazy_load
        self.lazy_load_file = lazy_load_file
        self.config = {}
        self.weight = None
        self.bias = None

    def load(self, weight_dict):
        if not self.lazy_load:
            self.weight = weight_dict[self.weight_name] if self.weight_name is not None else None
            self.bias = weight_dict[self.bias_name] if self.bias_name is not None else None

    def apply(self, input_tensor: torch.Tensor, *, scale=None, shift=None, out=None) -> torch.Tensor:
        return lib.ln_modulate(input_tensor, weight=self.weight, bias=self.bias, scale=scale, shift=shift, eps=self.eps, out=out)

    def state_dict(self, destination=None):
        if destination is None:
            destination = {}
        if self.weight is not None:
            destination[self.weight_name] = self.weight.cpu().detach().clone()
        if self.bias is not None:
            destination[self.bias_name] = self.bias.cpu().detach().clone()
        return destination


LN_WEIGHT_REGISTER["Default"] = LNWeightB200
LN_WEIGHT_REGISTER[MM_KEY] = LNWeightB200      # CogVideoX looks its LayerNorms up under the mm_type key (cogvideox/weights/transformers_weights.py:45-48)


@ATTN_WEIGHT_REGISTER(ATTN_KEY)
class FmhaWeightB200:
    """Varlen non-causal attention, same call signature as FlashAttn2Weight.apply (attn_weight.py:76-97).
    cu_seqlens are HOST-visible sequences or tensors; each segment is one kernel launch (Wan: one segment; Hunyuan: two)."""

    def __init__(self):
        self.config = {}
        self._cu_cache = {}

    def load(self, weight_dict):
        pass

    def set_config(self, config=None):
        if config is not None:
            self.config = config

    def to_cpu(self, non_blocking=False):
        pass

    def to_cuda(self, non_blocking=False):
        pass

    def state_dict(self, destination=None):
        return {} if destination is None else destination

    def _bounds(self, cu, total):
        """Host copy of a cu_seqlens argument.  A CUDA tensor costs one device->host read the first time it is seen; the result is cached per
        tensor object (+ `_version`), so the reference infer classes, which pass the same cu_seqlens every block, do not sync per call."""
        if cu is None:
            return [0, total]
        if isinstance(cu, torch.Tensor):
            ent = self._cu_cache.get(id(cu))
            if ent is not None and ent[0] is cu and ent[1] == cu._version:
                return ent[2]
            b = [int(v) for v in cu.tolist()]
            if len(self._cu_cache) >= 8:
                self._cu_cache.pop(next(iter(self._cu_cache)))
            self._cu_cache[id(cu)] = (cu, cu._version, b)
            return b
        return [int(v) for v in cu]

    def apply(self, q, k, v, cu_seqlens_q=None, cu_seqlens_kv=None, max_seqlen_q=None, max_seqlen_kv=None,
              model_cls=None, mask_map=None):
        bq = self._bounds(cu_seqlens_q, q.shape[0])
        bk = self._bounds(cu_seqlens_kv, k.shape[0])
        if len(bq) != len(bk):
            raise lib.B200Error("b200_fmha: cu_seqlens_q and cu_seqlens_kv must have the same number of segments")
        out = torch.empty((q.shape[0], q.shape[1], q.shape[2]), dtype=torch.bfloat16, device=q.device)
        covered = 0
        for i in range(len(bq) - 1):
            if bq[i + 1] > bq[i] and bk[i + 1] > bk[i]:
                if bq[i] > covered:
                    out[covered:bq[i]].zero_()                     # rows no segment attends from: defined (zero) output, like flash-attn's padded rows
                lib.fmha(q[bq[i]:bq[i + 1]], k[bk[i]:bk[i + 1]], v[bk[i]:bk[i + 1]], out=out[bq[i]:bq[i + 1]])
                covered = bq[i + 1]
        if covered < q.shape[0]:
            out[covered:].zero_()
        rows = q.shape[0] if max_seqlen_q is None else max_seqlen_q
        return out.reshape(rows, -1)


@TENSOR_REGISTER("Default")
class DefaultTensor(_WeightOp):
    _attrs = ("tensor",)

    def __init__(self, tensor_name, lazy_load=False, lazy_load_file=None):
        self.tensor_name = tensor_name
        self.lazy_load = lazy_load
        self.lazy_load_file = lazy_load_file
        self.tensor = None

    def load(self, weight_dict):
        if not self.lazy_load:
            self.tensor = weight_dict[self.tensor_name]

    def state_dict(self, destination=None):
        if destination is None:
            destination = {}
        destination[self.tensor_name] = self.tensor.cpu().detach().clone()
        return destination
