"""Wan DiT forward = pre_infer -> block stack -> post_infer, cond/uncond + CFG combine, with the reference's call order
(lightx2v/models/networks/wan/model.py:197-226) and the reference's pre/post math on the B200 kernels:

  WanPreInfer.infer   lightx2v/models/networks/wan/infer/pre_infer.py:29-120
      patch-embed conv3d k = s = (1,2,2)  == a [S, C*4] x [C*4, D] GEMM on the unfolded patches  (conv3d.py:40-50)
      sinusoidal t-embedding in fp64 -> bf16 (utils.py:161-172) -> time MLP -> 6-way projection
      text MLP with GELU(tanh);  i2v: CLIP MLP (LN -> Linear -> GELU(erf) -> Linear -> LN), concat
  WanPostInfer.infer  lightx2v/models/networks/wan/infer/post_infer.py:15-50
      LN -> (1+scale), shift from head.modulation + embed -> Linear(D -> 64) -> unpatchify -> fp32
Weights are a flat dict with the checkpoint's key names (patch_embedding.*, text_embedding.{0,2}.*, time_embedding.{0,2}.*,
time_projection.1.*, img_emb.proj.{0,1,3,4}.*, head.head.*, head.modulation, blocks.*)."""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .. import lib
from .wan_infer import WanTransformerInfer
from .registry import MM_KEY
from .wan_weights import WanTransformerWeights


def sinusoidal_embedding_1d(dim: int, position: torch.Tensor) -> torch.Tensor:
    """utils.py:161-172 (BF16 mode: fp64 math, one rounding to bf16)."""
    half = dim // 2
    position = position.type(torch.float64)
    sinusoid = torch.outer(position, torch.pow(10000, -torch.arange(half).to(position).div(half)))
    return torch.cat([torch.cos(sinusoid), torch.sin(sinusoid)], dim=1).to(torch.bfloat16)


def rope_params(max_seq_len: int, dim: int, theta: float = 10000.0) -> torch.Tensor:
    """utils.py:151-158."""
    freqs = torch.outer(torch.arange(max_seq_len), 1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim)))
    return torch.polar(torch.ones_like(freqs), freqs)


class WanPreInfer:
    def __init__(self, config):
        d = config["dim"] // config["num_heads"]
        self.task = config["task"]
        self.freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)), rope_params(1024, 2 * (d // 6))], dim=1)
        self.freq_dim = config["freq_dim"]
        self.dim = config["dim"]
        self.text_len = config["text_len"]
        self.scheduler = None
        self._t_table = None
        self._ctx_cache = {}
        self.use_static_t_embed = False      # host/wan_graph.py: read the embedding row from the scheduler's static buffer (CUDA-graph replay)

    def set_scheduler(self, scheduler):
        self.scheduler = scheduler
        self._t_table = None

    def _embed_context(self, W, context, clip_fea):
        """Text MLP (+ CLIP MLP for i2v) of pre_infer.py:89-111.  The prompt embeddings do not change between denoise steps, so the result is
        computed once per (context, clip) tensor pair and the SAME tensor object is handed to the blocks every step - which is what lets
        their per-context K/V caches hit.  Entries hold references to their inputs (no address aliasing) and check `_version`."""
        key = (id(context), None if clip_fea is None else id(clip_fea))
        ent = self._ctx_cache.get(key)
        if ent is not None and ent[0] is context and ent[1] == context._version and ent[2] is clip_fea and (clip_fea is None or ent[3] == clip_fea._version):
            return ent[4]
        ctx = context
        if ctx.shape[0] < self.text_len:
            ctx = torch.cat([ctx, ctx.new_zeros(self.text_len - ctx.shape[0], ctx.shape[1])])
        out = lib.gemm_bf16(ctx.contiguous(), W["text_embedding.0.weight"], W["text_embedding.0.bias"], epilogue=lib.EPI_BIAS_GELU)
        res = lib.gemm_bf16(out, W["text_embedding.2.weight"], W["text_embedding.2.bias"])
        if clip_fea is not None:
            c = lib.ln_modulate(clip_fea.contiguous(), weight=W["img_emb.proj.0.weight"], bias=W["img_emb.proj.0.bias"], eps=1e-6)
            c = lib.gemm_bf16(c, W["img_emb.proj.1.weight"], W["img_emb.proj.1.bias"])
            c = F.gelu(c, approximate="none")
            c = lib.gemm_bf16(c, W["img_emb.proj.3.weight"], W["img_emb.proj.3.bias"])
            c = lib.ln_modulate(c, weight=W["img_emb.proj.4.weight"], bias=W["img_emb.proj.4.bias"], eps=1e-6)
            res = torch.cat([c, res], dim=0)
        if len(self._ctx_cache) >= 4:
            self._ctx_cache.pop(next(iter(self._ctx_cache)))
        self._ctx_cache[key] = (context, context._version, clip_fea, None if clip_fea is None else clip_fea._version, res)
        return res

    def _t_embedding(self, device) -> torch.Tensor:
        """Sinusoidal embedding of the current timestep (pre_infer.py:56-58).  The fp64 table for ALL timesteps of the schedule is built
        once on the host (same element-wise math as the per-step call of the reference) and indexed on the device afterwards, so a
        denoise step has no device->host synchronisation."""
        if self.use_static_t_embed:
            return self.scheduler.t_embed_static
        ts = self.scheduler.timesteps
        if self._t_table is None or self._t_table[0] is not ts or self._t_table[1] != ts._version:      # keyed on the tensor object (kept alive here)
            self._t_table = (ts, ts._version, sinusoidal_embedding_1d(self.freq_dim, ts.flatten().cpu()).to(device))
        i = self.scheduler.step_index
        return self._t_table[2][i:i + 1]

    def infer(self, W: Dict[str, torch.Tensor], inputs, positive: bool):
        x = self.scheduler.latents                                       # [C, F, H, W] (bf16 after step_pre)
        context = inputs["text_encoder_output"]["context" if positive else "context_null"]
        if isinstance(context, (list, tuple)):                           # the reference passes a one-element list (pre_infer.py:90 stacks it)
            context = context[0]
        if self.task == "i2v":
            clip_fea = inputs["image_encoder_output"]["clip_encoder_out"]
            x = torch.cat([x, inputs["image_encoder_output"]["vae_encode_out"].to(x.dtype)], dim=0)
        C, Fr, H, Wd = x.shape
        gh, gw = H // 2, Wd // 2
        grid_sizes = torch.tensor([[Fr, gh, gw]], dtype=torch.long)
        # patch embedding as a GEMM: token (f, y, x) <- patch x[:, f, 2y:2y+2, 2x:2x+2] flattened in (c, dy, dx) order
        patches = x.to(torch.bfloat16).view(C, Fr, gh, 2, gw, 2).permute(1, 2, 4, 0, 3, 5).reshape(Fr * gh * gw, C * 4).contiguous()
        pw = W["patch_embedding.weight"].reshape(self.dim, C * 4)
        xs = lib.gemm_bf16(patches, pw, W["patch_embedding.bias"])
        seq_lens = torch.tensor([xs.shape[0]], dtype=torch.long)

        embed = self._t_embedding(xs.device)
        embed = lib.gemm_bf16(embed, W["time_embedding.0.weight"], W["time_embedding.0.bias"])
        embed = F.silu(embed)
        embed = lib.gemm_bf16(embed, W["time_embedding.2.weight"], W["time_embedding.2.bias"])
        embed0 = lib.gemm_bf16(F.silu(embed), W["time_projection.1.weight"], W["time_projection.1.bias"]).unflatten(1, (6, self.dim))

        context = self._embed_context(W, context, clip_fea if self.task == "i2v" else None)
        return embed, grid_sizes, (xs, embed0.squeeze(0), seq_lens, self.freqs, context)


class WanPostInfer:
    def __init__(self, config):
        self.out_dim = config["out_dim"]
        self.patch_size = (1, 2, 2)
        self.scheduler = None

    def set_scheduler(self, scheduler):
        self.scheduler = scheduler

    def infer(self, W, x, e, grid_sizes):
        mod = (W["head.modulation"] + e.unsqueeze(1)).reshape(2, -1)       # post_infer.py:16-18: [shift, scale]
        x = lib.ln_modulate(x, scale=mod[1], shift=mod[0], eps=1e-6)
        x = lib.gemm_bf16(x, W["head.head.weight"], W["head.head.bias"])
        return [u.float() for u in self.unpatchify(x, grid_sizes)]

    def unpatchify(self, x, grid_sizes):
        """post_infer.py:41-50."""
        c = self.out_dim
        out = []
        for u, v in zip(x.unsqueeze(0), grid_sizes.tolist()):
            u = u[: math.prod(v)].view(*v, *self.patch_size, c)
            u = torch.einsum("fhwpqrc->cfphqwr", u)
            out.append(u.reshape(c, *[i * j for i, j in zip(v, self.patch_size)]))
        return out


class WanModel:
    """`WanModel(model_path, config, device)` like the reference (lightx2v/models/networks/wan/model.py:33-59): loads every
    `*.safetensors` under `model_path` (or `model_path/original/`, model.py:79-95) as bf16 onto `device` and builds the weight tree and the
    pre / transformer / post infer objects; `config["feature_caching"]` ("NoCaching" | "Tea") and `config["model_cls"]` ("wan2.1_causvid")
    select the transformer infer class like `_init_infer_class` (:61-75; the CausVid subclass is wan/causvid/model.py);
    `config["parallel_attn_type"] == "ulysses"` installs the sequence-parallel hooks (:52-58).  Weights stay resident on the GPU
    (180 GB HBM3e: the reference's cpu_offload managers are not needed).  `WanModel.from_weight_dict(config, weight_dict)` builds the
    same object from an in-memory checkpoint-named dict (synthetic weights in the tests and bench.py)."""

    def __init__(self, model_path, config, device):
        self.model_path = model_path
        self.config = config
        self.device = torch.device(device) if device is not None else None
        self._init_infer_class()
        self._init_weights()
        self._init_infer()
        if config.get("parallel_attn_type"):
            self._init_parallel(config["parallel_attn_type"])

    @classmethod
    def from_weight_dict(cls, config, weight_dict: Dict[str, torch.Tensor], device=None):
        self = cls.__new__(cls)
        self.model_path = None
        self.config = config
        self.device = torch.device(device) if device is not None else None
        self._init_infer_class()
        self._init_weights(weight_dict)
        self._init_infer()
        if config.get("parallel_attn_type"):
            self._init_parallel(config["parallel_attn_type"])
        return self

    # ------------------------------------------------------------------ construction (model.py:61-140)
    def _init_infer_class(self):
        self.pre_infer_class = WanPreInfer
        self.post_infer_class = WanPostInfer
        caching = self.config.get("feature_caching", "NoCaching")
        if self.config.get("model_cls") == "wan2.1_causvid":
            from .wan_causvid import WanTransformerInferCausVid
            self.transformer_infer_class = WanTransformerInferCausVid
        elif caching == "NoCaching":
            self.transformer_infer_class = WanTransformerInfer
        elif caching == "Tea":
            from .wan_teacache import WanTransformerInferTeaCaching
            self.transformer_infer_class = WanTransformerInferTeaCaching
        else:       # TaylorSeer / Ada / Custom change which steps run, not the step
            raise NotImplementedError(f"Unsupported feature_caching type: {caching}")

    def _load_ckpt(self) -> Dict[str, torch.Tensor]:
        import glob
        import os

        from safetensors import safe_open

        files = glob.glob(os.path.join(self.model_path, "*.safetensors")) or glob.glob(os.path.join(self.model_path, "original", "*.safetensors"))
        if not files:
            raise FileNotFoundError(f"No .safetensors files found in directory: {self.model_path}")
        out = {}
        quant = (self.config.get("mm_config") or {}).get("mm_type", MM_KEY) != MM_KEY          # pre-quantised checkpoints keep their dtypes
        for fp in sorted(files):
            with safe_open(fp, framework="pt") as f:
                for key in f.keys():
                    t = f.get_tensor(key)
                    if t.is_floating_point() and t.dtype in (torch.float32, torch.float16) and not (quant and key.endswith(("weight_scale", "weight_global_scale"))):
                        t = t.to(torch.bfloat16)
                    out[key] = t.to(self.device)
        return out

    def _init_weights(self, weight_dict: Optional[Dict[str, torch.Tensor]] = None):
        self.W = self._load_ckpt() if weight_dict is None else weight_dict
        self.transformer_weights = WanTransformerWeights(self.config)
        self.transformer_weights.load(self.W)
        self.pre_weight = self.post_weight = self.W        # the reference's WanPreWeights / WanPostWeights: here the flat dict itself

    def _init_infer(self):
        self.pre_infer = self.pre_infer_class(self.config)
        self.post_infer = self.post_infer_class(self.config)
        self.transformer_infer = self.transformer_infer_class(self.config)
        self.scheduler = None
        self.pre_process = None      # sequence-parallel hooks (host/ulysses.py)
        self.post_process = None
        self.cfg_parallel = None     # (branch_is_cond, cond_src_rank, uncond_src_rank, group) - host/ulysses.py:parallelize_wan_cfg

    def _init_parallel(self, kind: str):
        from . import ulysses

        if kind != "ulysses":        # ring attention is out of scope (superseded by Ulysses on NVSwitch)
            raise Exception("Unsuppotred parallel_attn_type")
        ts = self.config["target_shape"]
        total = ts[1] * (ts[2] // 2) * (ts[3] // 2)
        ulysses.parallelize_wan_fused(self, total)

    def set_scheduler(self, scheduler):
        """model.py:180-185."""
        self.scheduler = scheduler
        self.pre_infer.set_scheduler(scheduler)
        self.post_infer.set_scheduler(scheduler)
        self.transformer_infer.set_scheduler(scheduler)

    def to_cpu(self):
        self.transformer_weights.to_cpu()

    def to_cuda(self):
        self.transformer_weights.to_cuda()
        self.transformer_infer.clear_weight_caches()

    def _forward(self, inputs, positive: bool):
        embed, grid_sizes, (x, embed0, seq_lens, freqs, context) = self.pre_infer.infer(self.W, inputs, positive)
        if self.pre_process is not None:
            x = self.pre_process(x)
        x = self.transformer_infer.infer(self.transformer_weights, grid_sizes, embed, x, embed0, seq_lens, freqs, context)
        if self.post_process is not None:
            x = self.post_process(x)
        return self.post_infer.infer(self.W, x, embed, grid_sizes)[0]

    @torch.no_grad()
    def infer(self, inputs):
        """model.py:197-226: cond pass, then (enable_cfg) uncond pass and `uncond + g * (cond - uncond)` in fp32."""
        if self.cfg_parallel is not None and self.config.get("enable_cfg", False):
            # CFG-parallel: the conditional and unconditional passes of model.py:203-218 are independent, so one half
            # of the ranks runs each and the two 19 MB predictions are exchanged; same kernels on the same inputs -> same values.
            import torch.distributed as dist

            is_cond, cond_src, uncond_src, group = self.cfg_parallel
            mine = self._forward(inputs, is_cond).contiguous()
            world = dist.get_world_size(group)
            # output = the inputs concatenated along dim 0 (the layout every backend accepts; gloo rejects the stacked form)
            both = torch.empty((world * mine.shape[0],) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
            dist.all_gather_into_tensor(both, mine, group=group)
            both = both.view((world,) + tuple(mine.shape))
            cond, uncond = both[cond_src], both[uncond_src]
            self.scheduler.noise_pred = uncond + self.config["sample_guide_scale"] * (cond - uncond)
            return
        cond = self._forward(inputs, True)
        self.scheduler.noise_pred = cond
        if self.config.get("enable_cfg", False):
            uncond = self._forward(inputs, False)
            self.scheduler.noise_pred = uncond + self.config["sample_guide_scale"] * (cond - uncond)
