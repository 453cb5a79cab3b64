#!/usr/bin/env python
"""Generate golden fixtures from the REAL LightX2V reference (runs only in the build container).

Imports `/root/reference/lightx2v` under the two shims of SURVEY.md §8c (patch torch.cuda.get_device_capability; drop
pin_memory), builds the reference's own weight tree (`WanTransformerAttentionBlock`) + infer class
(`WanTransformerInfer`) on seeded synthetic weights, runs them on CPU in DTYPE=BF16 mode with `torch_sdpa`
attention, and stores inputs and outputs as safetensors under tests/golden/.  The fixtures pin oracle/wan_oracle.py
(tests/test_oracle_golden.py) and are also compared with the CUDA path on the GPU box (tests/test_gpu_block.py).

    python oracle/gen_golden.py            # writes tests/golden/*.safetensors
"""
import os
import sys
import types

os.environ["DTYPE"] = "BF16"
os.environ.setdefault("ENABLE_GRAPH_MODE", "false")

import torch  # noqa: E402

REF = os.environ.get("LIGHTX2V_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def install_shims():
    """Shim 1: attn_weight.py:26 / sage_attn2.py:3 call torch.cuda.get_device_capability(0) at import.
    Shim 2: every op's load() allocates torch.empty(..., pin_memory=True) (mm_weight.py:77, rms_norm_weight.py:23...)."""
    torch.cuda.get_device_capability = lambda *a, **k: (10, 0)
    _empty = torch.empty

    def empty_nopin(*a, **k):
        k.pop("pin_memory", None)
        return _empty(*a, **k)

    torch.empty = empty_nopin
    sys.path.insert(0, REF)


class Cfg(dict):
    """EasyDict stand-in (easydict is not installed): dict with attribute access."""

    __getattr__ = dict.__getitem__


def ref_config(dim, num_heads, ffn_dim, num_layers, task="t2v"):
    return Cfg(
        task=task, num_layers=num_layers, num_heads=num_heads, dim=dim, ffn_dim=ffn_dim, cpu_offload=False,
        mm_config={}, do_mm_calib=False, self_attn_1_type="torch_sdpa", cross_attn_1_type="torch_sdpa",
        cross_attn_2_type="torch_sdpa", model_cls="wan2.1", freq_dim=256, text_len=512, in_dim=16, out_dim=16,
        enable_cfg=True, attention_type="torch_sdpa",
    )


def run_reference_blocks(W, cfg, x, embed0, grid_sizes, freqs, context):
    from lightx2v.models.networks.wan.infer.transformer_infer import WanTransformerInfer
    from lightx2v.models.networks.wan.weights.transformer_weights import WanTransformerWeights

    weights = WanTransformerWeights(cfg)
    weights.load(W)
    infer = WanTransformerInfer(cfg)
    seq_lens = torch.tensor([x.shape[0]], dtype=torch.long)
    gs = torch.tensor([list(grid_sizes)], dtype=torch.long)
    return infer.infer(weights, gs, None, x, embed0, seq_lens, freqs, context)


def main():
    install_shims()
    from safetensors.torch import save_file

    import lightx2v.common.ops  # noqa: F401  registers MM/ATTN/RMS/LN/TENSOR op classes (common/ops/__init__.py)
    from lightx2v.common.ops import mm, attn, norm, tensor, conv  # noqa: F401

    from oracle import wan_oracle as O

    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)

    # ---- fixture 1: two Wan-1.3B-shaped blocks (D 1536, 12 heads, F 8960), grid 3x8x10 -> 240 tokens, t2v
    # ---- fixture 2: one block, i2v (257 CLIP rows + 512 text rows), grid 2x6x9 -> 108 tokens
    for name, task, L, grid in (("wan13b_t2v_2blocks", "t2v", 2, (3, 8, 10)), ("wan13b_i2v_1block", "i2v", 1, (2, 6, 9))):
        dim, heads, ffn = 1536, 12, 8960
        S = grid[0] * grid[1] * grid[2]
        cfg = ref_config(dim, heads, ffn, L, task)
        W = O.synth_block_weights(L, dim, ffn, task=task, seed=42)
        x, embed0, context = O.synth_block_inputs(S, dim, task=task, seed=7)
        freqs = O.wan_freqs_table(dim // heads)
        x_in = x.clone()
        out = run_reference_blocks(W, cfg, x, embed0, grid, freqs, context)
        # per-function probes from the reference's own methods on block 0 (fresh x)
        from lightx2v.models.networks.wan.infer.transformer_infer import WanTransformerInfer
        from lightx2v.models.networks.wan.infer.utils import apply_rotary_emb, compute_freqs
        from lightx2v.models.networks.wan.weights.transformer_weights import WanTransformerWeights

        weights = WanTransformerWeights(cfg)
        weights.load(W)
        infer = WanTransformerInfer(cfg)
        blk = weights.blocks[0]
        xs = x_in.clone()
        gs = torch.tensor([list(grid)], dtype=torch.long)
        seq_lens = torch.tensor([S], dtype=torch.long)
        sh, sc, ga, csh, csc, cga = infer.infer_modulation(blk.compute_phases[0], embed0)
        y_out = infer.infer_self_attn(blk.compute_phases[1], gs, xs, seq_lens, freqs, sh, sc)
        xs2, attn_out = infer.infer_cross_attn(blk.compute_phases[2], xs, context, y_out, ga)
        x_after_cross = xs2.clone()
        y_ffn = infer.infer_ffn(blk.compute_phases[3], xs2, attn_out, csh, csc)
        freqs_i = compute_freqs(dim // heads // 2, gs, freqs)
        qprobe = torch.randn(S, heads, dim // heads, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16)
        rope_out = apply_rotary_emb(qprobe, freqs_i)
        tensors = {
            "x_in": x_in, "embed0": embed0, "context": context, "x_out": out,
            "probe.self_attn_y": y_out, "probe.x_after_cross": x_after_cross, "probe.cross_attn_out": attn_out,
            "probe.ffn_y": y_ffn, "probe.rope_in": qprobe, "probe.rope_out": rope_out,
            "grid": torch.tensor(grid, dtype=torch.int64),
        }
        meta = {"dim": str(dim), "heads": str(heads), "ffn": str(ffn), "layers": str(L), "task": task, "weights_seed": "42",
                "inputs_seed": "7", "generator": "oracle/gen_golden.py", "reference": "ModelTC/lightx2v@0591c35e",
                "attn": "torch_sdpa", "dtype_mode": "BF16"}
        save_file({k: v.contiguous() for k, v in tensors.items()}, os.path.join(GOLD, name + ".safetensors"), metadata=meta)
        print(name, "x_out absmax", float(out.float().abs().max()), "bytes",
              os.path.getsize(os.path.join(GOLD, name + ".safetensors")))


def gen_scheduler_fixture():
    """Real WanScheduler (lightx2v/models/schedulers/wan/scheduler.py) driven for 8 of 20 steps on CPU with seeded pseudo
    model outputs; pins lightx2v_b200/host/wan_scheduler.py (tests/test_host_cpu.py)."""
    from safetensors.torch import save_file

    from lightx2v.models.schedulers.wan.scheduler import WanScheduler

    cfg = Cfg(infer_steps=20, target_video_length=17, sample_shift=5.0, seed=42, task="t2v", target_shape=(16, 3, 8, 8), patch_size=(1, 2, 2))
    sch = WanScheduler(cfg)
    sch.device = torch.device("cpu")
    sch.prepare()
    g = torch.Generator().manual_seed(11)
    tensors = {"latents_0": sch.latents.clone(), "timesteps": sch.timesteps.clone(), "sigmas": sch.sigmas.clone()}
    for i in range(8):
        sch.step_pre(i)
        tensors[f"latents_pre_{i}"] = sch.latents.clone()
        sch.noise_pred = torch.randn(sch.latents.shape, generator=g)
        tensors[f"noise_pred_{i}"] = sch.noise_pred.clone()
        sch.step_post()
        tensors[f"latents_post_{i}"] = sch.latents.clone()
    save_file({k: v.contiguous() for k, v in tensors.items()}, os.path.join(GOLD, "wan_scheduler_unipc.safetensors"),
              metadata={"infer_steps": "20", "sample_shift": "5.0", "seed": "42", "generator": "oracle/gen_golden.py:gen_scheduler_fixture"})
    print("wan_scheduler_unipc", os.path.getsize(os.path.join(GOLD, "wan_scheduler_unipc.safetensors")))


def gen_vae_fixture():
    """Real WanVAE_ (lightx2v/models/video_encoders/hf/wan/vae.py) decoder, seeded synthetic weights loaded through its own
    state_dict, decode of a [16, 3, 8, 8] latent -> [1, 3, 9, 64, 64]; pins oracle/vae_oracle.py."""
    from safetensors.torch import save_file

    from lightx2v.models.video_encoders.hf.wan.vae import WanVAE_

    from oracle import vae_oracle as V

    W = V.synth_vae_weights(seed=0)
    model = WanVAE_(dim=96, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[], temperal_downsample=[False, True, True], dropout=0.0).eval()
    res = model.load_state_dict(W, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.startswith("encoder.") or k.startswith("conv1.") for k in res.missing_keys), res.missing_keys
    g = torch.Generator().manual_seed(5)
    zs = torch.randn(16, 3, 8, 8, generator=g)
    scale = [torch.tensor(V.MEAN), 1.0 / torch.tensor(V.STD)]
    with torch.no_grad():
        out = model.decode(zs.unsqueeze(0), scale).float().clamp_(-1, 1)
    save_file({"zs": zs, "images": out.contiguous()}, os.path.join(GOLD, "wan_vae_decode_small.safetensors"),
              metadata={"weights_seed": "0", "generator": "oracle/gen_golden.py:gen_vae_fixture", "reference": "ModelTC/lightx2v@0591c35e"})
    print("wan_vae_decode_small", tuple(out.shape), "absmax", float(out.abs().max()), "frac clamped", float((out.abs() >= 1).float().mean()))


def gen_hunyuan_fixture():
    """Real HunyuanTransformerInfer (lightx2v/models/networks/hunyuan/infer/transformer_infer.py) on one double-stream and one
    single-stream block at the model's true width (3072, 24 heads, MLP 12288), 96 image + 32 text tokens, torch_sdpa attention with
    a single segment (the reference's torch_sdpa op ignores cu_seqlens; two-segment varlen is checked on the GPU against flash-attn)."""
    from safetensors.torch import save_file

    import lightx2v.common.ops  # noqa: F401
    from lightx2v.models.networks.hunyuan.infer.transformer_infer import HunyuanTransformerInfer
    from lightx2v.models.networks.hunyuan.weights.transformer_weights import HunyuanTransformerDoubleBlock, HunyuanTransformerSingleBlock

    from oracle import hunyuan_oracle as HO

    hidden, mlp, heads = 3072, 12288, 24
    cfg = Cfg(cpu_offload=False, do_mm_calib=False, mm_config={}, attention_type="torch_sdpa", task="t2v")
    W = HO.synth_weights(1, 1, hidden, mlp, seed=42)
    img, txt, vec, cu, freqs = HO.synth_inputs(96, 32, 32, hidden, seed=7)
    dbl, sgl = HunyuanTransformerDoubleBlock(0, cfg), HunyuanTransformerSingleBlock(0, cfg)
    dbl.load(W)
    sgl.load(W)
    infer = HunyuanTransformerInfer(cfg)
    cu_t = torch.tensor(cu, dtype=torch.int32)
    max_len = img.shape[0] + txt.shape[0]
    img1, txt1 = infer.infer_double_block(dbl, img.clone(), txt.clone(), vec, cu_t, max_len, freqs, None, None)
    x = torch.cat((img1, txt1), 0)
    x2 = infer.infer_single_block(sgl, x, vec, txt.shape[0], cu_t, max_len, freqs, None, None)
    save_file({"img": img, "txt": txt, "vec": vec, "cos": freqs[0], "sin": freqs[1], "img_after_double": img1.contiguous(),
               "txt_after_double": txt1.contiguous(), "x_after_single": x2.contiguous()},
              os.path.join(GOLD, "hunyuan_blocks_small.safetensors"),
              metadata={"hidden": str(hidden), "mlp": str(mlp), "heads": str(heads), "weights_seed": "42", "inputs_seed": "7", "txt_valid": "32",
                        "generator": "oracle/gen_golden.py:gen_hunyuan_fixture", "reference": "ModelTC/lightx2v@0591c35e"})
    print("hunyuan_blocks_small", float(x2.float().abs().max()), os.path.getsize(os.path.join(GOLD, "hunyuan_blocks_small.safetensors")))


def install_diffusers_shim():
    """`diffusers` is not installed in this image.  The reference's HunyuanVideo VAE files import it for (a) config / model mixins
    and (b) ONE compute layer, `Attention` (mid-block, `_from_deprecated_attn_block=True`).  This shim supplies inert mixins and a
    restatement of that layer's published algorithm (diffusers 0.31 `AttnProcessor2_0`): GroupNorm -> q/k/v Linear -> SDPA with the
    additive mask -> to_out -> + residual -> / rescale.  Everything else on the decode path is the reference's own code."""
    import inspect
    from dataclasses import dataclass  # noqa: F401

    import torch.nn as nn
    import torch.nn.functional as F

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    d = mod("diffusers")
    cu = mod("diffusers.configuration_utils")
    ld = mod("diffusers.loaders")
    ut = mod("diffusers.utils")
    au = mod("diffusers.utils.accelerate_utils")
    tu = mod("diffusers.utils.torch_utils")
    md = mod("diffusers.models")
    ap = mod("diffusers.models.attention_processor")
    mo = mod("diffusers.models.modeling_outputs")
    mu = mod("diffusers.models.modeling_utils")
    ac = mod("diffusers.models.activations")
    nm = mod("diffusers.models.normalization")
    d.models = md

    class ConfigMixin:
        pass

    def register_to_config(init):
        def wrapped(self, *a, **k):
            sig = inspect.signature(init)
            bound = sig.bind(self, *a, **k)
            bound.apply_defaults()
            self.config = Cfg({n: v for n, v in bound.arguments.items() if n != "self"})
            init(self, *a, **k)
        return wrapped

    cu.ConfigMixin, cu.register_to_config = ConfigMixin, register_to_config
    ld.FromOriginalVAEMixin = type("FromOriginalVAEMixin", (), {})
    au.apply_forward_hook = lambda f: f

    class BaseOutput:
        pass

    class _Log:
        def warn(self, *a, **k):
            pass
        warning = info = debug = warn

    ut.BaseOutput = BaseOutput
    ut.is_torch_version = lambda op, v: True
    ut.logging = types.SimpleNamespace(get_logger=lambda name: _Log())
    tu.randn_tensor = lambda shape, generator=None, device=None, dtype=None: torch.randn(shape, generator=generator, device=device, dtype=dtype)

    class Attention(nn.Module):
        def __init__(self, query_dim, heads=8, dim_head=64, rescale_output_factor=1.0, eps=1e-5, norm_num_groups=None, spatial_norm_dim=None,
                     residual_connection=False, bias=False, upcast_softmax=False, _from_deprecated_attn_block=False, **kw):
            super().__init__()
            assert heads == 1 and spatial_norm_dim is None and _from_deprecated_attn_block
            inner = heads * dim_head
            self.group_norm = nn.GroupNorm(num_channels=query_dim, num_groups=norm_num_groups, eps=eps, affine=True)
            self.to_q, self.to_k, self.to_v = nn.Linear(query_dim, inner, bias=bias), nn.Linear(query_dim, inner, bias=bias), nn.Linear(query_dim, inner, bias=bias)
            self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(0.0)])
            self.residual_connection, self.rescale_output_factor = residual_connection, rescale_output_factor

        def forward(self, hidden_states, temb=None, attention_mask=None, **kw):
            residual = hidden_states
            h = self.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
            q, k, v = self.to_q(h).unsqueeze(1), self.to_k(h).unsqueeze(1), self.to_v(h).unsqueeze(1)
            m = attention_mask.unsqueeze(1) if attention_mask is not None else None
            o = F.scaled_dot_product_attention(q, k, v, attn_mask=m, dropout_p=0.0, is_causal=False).squeeze(1)
            o = self.to_out[1](self.to_out[0](o))
            if self.residual_connection:
                o = o + residual
            return o / self.rescale_output_factor

    ap.Attention = Attention
    for n in ("AttentionProcessor", "AttnAddedKVProcessor", "AttnProcessor", "SpatialNorm"):
        setattr(ap, n, type(n, (nn.Module,), {}))
    ap.ADDED_KV_ATTENTION_PROCESSORS, ap.CROSS_ATTENTION_PROCESSORS = (), ()
    mo.AutoencoderKLOutput = type("AutoencoderKLOutput", (BaseOutput,), {})
    mu.ModelMixin = type("ModelMixin", (nn.Module,), {})
    ac.get_activation = lambda name: {"swish": nn.SiLU, "silu": nn.SiLU}[name]()
    nm.AdaGroupNorm = type("AdaGroupNorm", (nn.Module,), {})
    nm.RMSNorm = type("RMSNorm", (nn.Module,), {})


def gen_hunyuan_vae_fixture():
    """Real AutoencoderKLCausal3D (lightx2v/models/video_encoders/hf/autoencoder_kl_causal_3d/) decode with tiling enabled, exactly
    the calls of VideoEncoderKLCausal3DModel.decode (model.py:33-44), fp32 on CPU, narrow widths (64, 64, 128, 128), sample_size 64 /
    sample_tsize 16 so that a [16, 6, 12, 10] latent exercises temporal tiles (2), spatial tiles (2 x 2) and all three blends."""
    from safetensors.torch import save_file

    install_diffusers_shim()
    from lightx2v.models.video_encoders.hf.autoencoder_kl_causal_3d.autoencoder_kl_causal_3d import AutoencoderKLCausal3D

    from oracle import hunyuan_vae_oracle as HV

    cfg = dict(HV.HUNYUAN_VAE_CFG, block_out_channels=(64, 64, 128, 128), sample_size=64, sample_tsize=16)
    W = HV.synth_vae_weights(cfg, seed=3)
    model = AutoencoderKLCausal3D(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlockCausal3D",) * 4, up_block_types=("UpDecoderBlockCausal3D",) * 4,
                                  block_out_channels=cfg["block_out_channels"], layers_per_block=2, act_fn="silu", latent_channels=16, norm_num_groups=32,
                                  sample_size=cfg["sample_size"], sample_tsize=cfg["sample_tsize"], scaling_factor=cfg["scaling_factor"],
                                  time_compression_ratio=4, spatial_compression_ratio=8, mid_block_add_attention=True).eval()
    res = model.load_state_dict(W, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.startswith("encoder.") or k.startswith("quant_conv.") for k in res.missing_keys), res.missing_keys
    g = torch.Generator().manual_seed(11)
    latents = torch.randn(1, 16, 6, 12, 10, generator=g)
    with torch.no_grad():
        z = latents / model.config.scaling_factor
        model.enable_tiling()
        image = model.decode(z, return_dict=False, generator=None)[0]
        image = (image / 2 + 0.5).clamp(0, 1).float()
        one = model.decoder(model.post_quant_conv(z[:, :, :3, :8, :8]))          # a single un-tiled tile, before the clamp
    save_file({"latents": latents, "images": image.contiguous(), "tile_raw": one.contiguous()}, os.path.join(GOLD, "hunyuan_vae_decode_small.safetensors"),
              metadata={"weights_seed": "3", "block_out_channels": "64,64,128,128", "sample_size": "64", "sample_tsize": "16",
                        "generator": "oracle/gen_golden.py:gen_hunyuan_vae_fixture", "reference": "ModelTC/lightx2v@0591c35e"})
    print("hunyuan_vae_decode_small", tuple(image.shape), "mean", float(image.mean()), "frac clamped", float(((image <= 0) | (image >= 1)).float().mean()),
          "tile_raw absmax", float(one.abs().max()))


def gen_nvfp4_fixture():
    """The reference's own NVFP4 golden model (lightx2v_kernel/test/nvfp4_nvfp4/fake_quant.py, imported unmodified) on a seeded
    [200, 256] activation and [96, 256] weight: e2m1 grid values, e4m3 scale values, and the dequantised-matmul result the reference's
    GEMM test compares against (test_bench1.py:75-103); pins oracle/nvfp4_oracle.py."""
    from safetensors.torch import save_file

    sys.path.insert(0, os.path.join(REF, "lightx2v_kernel", "test", "nvfp4_nvfp4"))
    import fake_quant as FQ

    g = torch.Generator().manual_seed(21)
    a = (torch.randn(200, 256, generator=g) * 1.7).to(torch.bfloat16)
    b = (torch.randn(96, 256, generator=g) * 0.05).to(torch.bfloat16)
    a[5, 32:48] = 0                                                     # an all-zero group: scale 0, reciprocal guarded
    a[7, 0] = 40.0                                                      # an outlier: other groups get small scales
    bias = torch.randn(96, generator=g).to(torch.bfloat16)
    gs_a = (448.0 * 6.0 / a.float().abs().max()).to(torch.float32)
    gs_b = (448.0 * 6.0 / b.float().abs().max()).to(torch.float32)
    qa, sa = FQ.ref_nvfp4_quant(a.clone(), gs_a)
    qb, sb = FQ.ref_nvfp4_quant(b.clone(), gs_b)
    da = (qa.reshape(200, 16, 16) * (sa / gs_a).unsqueeze(-1)).reshape(200, 256)
    db = (qb.reshape(96, 16, 16) * (sb / gs_b).unsqueeze(-1)).reshape(96, 256)
    out = da @ db.t() + bias.float()
    save_file({"a": a, "b": b, "bias": bias, "gs_a": gs_a.reshape(1), "gs_b": gs_b.reshape(1), "qa": qa.contiguous(), "sa": sa.contiguous(),
               "qb": qb.contiguous(), "sb": sb.contiguous(), "out": out.contiguous()}, os.path.join(GOLD, "nvfp4_quant_small.safetensors"),
              metadata={"generator": "oracle/gen_golden.py:gen_nvfp4_fixture", "reference": "ModelTC/lightx2v@0591c35e"})
    print("nvfp4_quant_small", float(out.abs().max()), "zero-scale groups", int((sa == 0).sum()))


def gen_teacache_fixture():
    """Decision sequence of the REAL WanTransformerInferTeaCaching.calculate_should_calc (feature_caching/transformer_infer.py:30-82) on a
    seeded random walk of timestep embeddings, cond / uncond passes interleaved as WanModel.infer does, for both `use_ret_steps` modes
    with the coefficients of configs/caching/teacache/wan_t2v_1_3b_tea_480p.json; `.cuda()` is patched to the identity (CPU run)."""
    import json

    from safetensors.torch import save_file

    from lightx2v.models.networks.wan.infer.feature_caching.transformer_infer import WanTransformerInferTeaCaching

    torch.Tensor.cuda = lambda self, *a, **k: self
    ref_cfg = json.load(open(os.path.join(REF, "configs", "caching", "teacache", "wan_t2v_1_3b_tea_480p.json")))
    steps = 20
    g = torch.Generator().manual_seed(17)
    base = torch.randn(1, 1536, generator=g)
    drift = [base * (1.0 + 0.01 * i) + 0.012 * (1 + (i % 4)) * torch.randn(1, 1536, generator=g) for i in range(steps)]
    embeds = torch.stack([d.to(torch.bfloat16) for d in drift])                       # [steps, 1, 1536]
    embed0s = torch.stack([(d.repeat(1, 6) * 0.5).to(torch.bfloat16).view(6, 1536) for d in drift])
    out = {"embeds": embeds, "embed0s": embed0s}
    for mode in (True, False):
        cfg = Cfg(task="t2v", num_layers=1, num_heads=12, dim=1536, cpu_offload=False, infer_steps=steps, enable_cfg=True,
                  teacache_thresh=ref_cfg["teacache_thresh"], coefficients=ref_cfg["coefficients"], use_ret_steps=mode)
        ti = WanTransformerInferTeaCaching(cfg)
        rec = []
        for i in range(steps):
            for cond in (True, False):
                ti.infer_conditional = cond
                rec.append(bool(ti.calculate_should_calc(embeds[i], embed0s[i])))
                ti.cnt += 1
        out[f"decisions_ret{int(mode)}"] = torch.tensor(rec, dtype=torch.uint8)
    save_file(out, os.path.join(GOLD, "wan_teacache_decisions.safetensors"),
              metadata={"thresh": str(ref_cfg["teacache_thresh"]), "coefficients": json.dumps(ref_cfg["coefficients"]),
                        "generator": "oracle/gen_golden.py:gen_teacache_fixture", "reference": "ModelTC/lightx2v@0591c35e"})
    print("wan_teacache_decisions", {k: int(v.sum()) for k, v in out.items() if k.startswith("dec")}, "of", 2 * steps)


def gen_causvid_fixture():
    """Real WanTransformerInferCausVid (wan/infer/causvid/transformer_infer.py) - the autoregressive block variant with a self-attention KV
    cache - on two Wan-1.3B-width blocks, three chunks of one latent frame (6 x 8 = 48 tokens each) fed in order with kv_start / kv_end
    advancing, t2v, torch_sdpa.  Outputs per chunk + the final K cache of block 0 pin oracle.wan_oracle.infer_blocks_causvid."""
    from safetensors.torch import save_file

    import lightx2v.common.ops  # noqa: F401
    from lightx2v.models.networks.wan.infer.causvid.transformer_infer import WanTransformerInferCausVid
    from lightx2v.models.networks.wan.weights.transformer_weights import WanTransformerWeights

    from oracle import wan_oracle as O

    dim, heads, ffn, L = 1536, 12, 8960, 2
    frame_tokens, chunks = 48, 3
    cfg = ref_config(dim, heads, ffn, L, "t2v")
    cfg.update(num_frames=chunks, num_frame_per_block=1, frame_seq_length=frame_tokens, model_cls="wan2.1_causvid")
    W = O.synth_block_weights(L, dim, ffn, seed=42)
    weights = WanTransformerWeights(cfg)
    weights.load(W)
    infer = WanTransformerInferCausVid(cfg)
    infer._init_kv_cache(torch.bfloat16, "cpu")
    infer._init_crossattn_cache(torch.bfloat16, "cpu")
    freqs = O.wan_freqs_table(dim // heads)
    gs = torch.tensor([[1, 6, 8]], dtype=torch.long)
    tensors = {}
    context = None
    for c in range(chunks):
        x, embed0, ctx = O.synth_block_inputs(frame_tokens, dim, seed=100 + c)
        context = ctx if context is None else context                      # the prompt does not change between chunks
        tensors[f"x_in.{c}"], tensors[f"embed0.{c}"] = x.clone(), embed0
        out = infer.infer(weights, gs, None, x, embed0, torch.tensor([frame_tokens]), freqs, context, c * frame_tokens, (c + 1) * frame_tokens)
        tensors[f"x_out.{c}"] = out.clone()
    tensors["context"] = context
    tensors["k_cache.0"] = infer.kv_cache[0]["k"].reshape(chunks * frame_tokens, dim).clone()
    save_file({k: v.contiguous() for k, v in tensors.items()}, os.path.join(GOLD, "wan13b_causvid_2blocks.safetensors"),
              metadata={"dim": str(dim), "heads": str(heads), "ffn": str(ffn), "layers": str(L), "chunks": str(chunks), "frame_tokens": str(frame_tokens),
                        "grid": "1,6,8", "weights_seed": "42", "generator": "oracle/gen_golden.py:gen_causvid_fixture", "reference": "ModelTC/lightx2v@0591c35e"})
    print("wan13b_causvid_2blocks", [float(tensors[f"x_out.{c}"].float().abs().max()) for c in range(chunks)])


def gen_prepost_fixture():
    """Real WanPreInfer / WanPostInfer (wan/infer/pre_infer.py, post_infer.py) with the reference's own WanPreWeights / WanPostWeights on
    seeded weights at Wan-1.3B width, latent [16, 3, 8, 12] (t2v) and the i2v variant (36 input channels, 257 CLIP tokens); `.cuda()` is
    patched to the identity.  Pins oracle.wan_oracle.pre_infer / post_infer (A13)."""
    from safetensors.torch import save_file

    import lightx2v.common.ops  # noqa: F401
    from lightx2v.models.networks.wan.infer.post_infer import WanPostInfer
    from lightx2v.models.networks.wan.infer.pre_infer import WanPreInfer
    from lightx2v.models.networks.wan.weights.post_weights import WanPostWeights
    from lightx2v.models.networks.wan.weights.pre_weights import WanPreWeights

    from oracle import wan_oracle as O

    torch.Tensor.cuda = lambda self, *a, **k: self
    dim, heads = 1536, 12
    out = {}
    for task in ("t2v", "i2v"):
        cfg = ref_config(dim, heads, 8960, 1, task)
        cfg.update(in_dim=36 if task == "i2v" else 16, model_cls="wan2.1")
        W = O.synth_prepost_weights(dim, cfg["in_dim"], task, seed=13)
        pre_w, post_w = WanPreWeights(cfg), WanPostWeights(cfg)
        pre_w.load(W)
        post_w.load(W)
        g = torch.Generator().manual_seed(31)
        latents = torch.randn(16, 3, 8, 12, generator=g).to(torch.bfloat16)
        context = torch.randn(77, 4096, generator=g).to(torch.bfloat16)
        timesteps = torch.tensor([999.0, 937.5, 612.25], dtype=torch.float32)
        sched = types.SimpleNamespace(latents=latents, timesteps=timesteps, step_index=1, flag_df=False, seq_len=3 * 4 * 6)
        inputs = {"text_encoder_output": {"context": [context], "context_null": [context]}, "image_encoder_output": None}
        if task == "i2v":
            inputs["image_encoder_output"] = {"clip_encoder_out": torch.randn(257, 1280, generator=g).to(torch.bfloat16),
                                              "vae_encode_out": torch.randn(20, 3, 8, 12, generator=g).to(torch.bfloat16)}
        pre, post = WanPreInfer(cfg), WanPostInfer(cfg)
        pre.set_scheduler(sched)
        post.set_scheduler(sched)
        embed, grid_sizes, (x, embed0, seq_lens, freqs, ctx) = pre.infer(pre_w, inputs, True)
        x_blocks = torch.randn(x.shape, generator=torch.Generator().manual_seed(37)).to(torch.bfloat16)       # stand-in for the block stack's output
        noise = post.infer(post_w, x_blocks.clone(), embed, grid_sizes)[0]
        t = {"latents": latents, "context": context, "timesteps": timesteps, "embed": embed, "x": x, "embed0": embed0, "context_out": ctx,
             "x_blocks": x_blocks, "noise_pred": noise, "grid": grid_sizes[0]}
        if task == "i2v":
            t.update(inputs["image_encoder_output"])
        out.update({f"{task}.{k}": v.contiguous() for k, v in t.items()})
    save_file(out, os.path.join(GOLD, "wan13b_prepost.safetensors"),
              metadata={"dim": str(dim), "weights_seed": "13", "step_index": "1", "generator": "oracle/gen_golden.py:gen_prepost_fixture",
                        "reference": "ModelTC/lightx2v@0591c35e"})
    print("wan13b_prepost", {k: tuple(v.shape) for k, v in out.items() if k.endswith((".x", ".context_out", ".noise_pred"))})


def gen_distill_scheduler_fixture():
    """Real WanStepDistillScheduler (schedulers/wan/step_distill/scheduler.py) for its 4 steps on CPU with seeded pseudo model outputs; the
    re-noising draws from the default generator (:53), so the global seed is set before every step_post and recorded."""
    from safetensors.torch import save_file

    from lightx2v.models.schedulers.wan.step_distill.scheduler import WanStepDistillScheduler

    cfg = Cfg(infer_steps=4, target_video_length=17, sample_shift=5.0, seed=42, task="t2v", target_shape=(16, 3, 8, 8), patch_size=(1, 2, 2),
              denoising_step_list=[1000, 750, 500, 250])
    sch = WanStepDistillScheduler(cfg)
    sch.device = torch.device("cpu")
    sch.prepare(None)
    g = torch.Generator().manual_seed(19)
    tensors = {"latents_0": sch.latents.clone(), "timesteps": sch.timesteps.clone(), "sigmas": sch.sigmas.clone()}
    for i in range(4):
        sch.step_pre(i)
        sch.noise_pred = torch.randn(sch.latents.shape, generator=g)
        tensors[f"noise_pred_{i}"] = sch.noise_pred.clone()
        torch.manual_seed(1000 + i)
        sch.step_post()
        tensors[f"latents_post_{i}"] = sch.latents.clone()
    save_file({k: v.contiguous() for k, v in tensors.items()}, os.path.join(GOLD, "wan_scheduler_step_distill.safetensors"),
              metadata={"sample_shift": "5.0", "seed": "42", "step_seed_base": "1000", "generator": "oracle/gen_golden.py:gen_distill_scheduler_fixture"})
    print("wan_scheduler_step_distill", sch.timesteps.tolist(), sch.sigmas.tolist())


def gen_cogvideox_fixture():
    """REAL CogvideoxTransformerInfer + CogVideoXBlock weight classes (lightx2v/models/networks/cogvideox/infer/transformer_infer.py:45-145,
    weights/transformers_weights.py:30-77; mm / LN ops "Default") on CPU: 2 blocks, 12 heads x 64, ff 3072, 40 text + 3x6x8 = 144 video
    tokens.  Pins oracle/cogvideox_oracle.py (tests/test_oracle_golden.py) and is compared with the CUDA path (tests/test_gpu_cogvideox.py).
    The rotary table has get_3d_rotary_pos_embed's shape and pair structure with synthetic angles (diffusers is not in the image)."""
    from safetensors.torch import save_file

    import lightx2v.common.ops  # noqa: F401
    from lightx2v.common.ops import mm, norm  # noqa: F401
    from lightx2v.models.networks.cogvideox.infer.transformer_infer import CogvideoxTransformerInfer
    from lightx2v.models.networks.cogvideox.weights.transformers_weights import CogvideoxTransformerWeights

    from oracle import cogvideox_oracle as C

    layers, heads, hd, ff, Lt, grid = 2, 12, 64, 3072, 40, (3, 6, 8)
    dim, Li = heads * hd, grid[0] * grid[1] * grid[2]
    W = C.synth_weights(layers, dim, ff, hd, seed=42)
    hidden, enc, temb = C.synth_inputs(Lt, Li, dim, seed=7)
    rotary = C.rotary_table(*grid, head_dim=hd, seed=3)
    cfg = Cfg(num_layers=layers, transformer_num_layers=layers, transformer_num_attention_heads=heads, transformer_attention_head_dim=hd, text_len=Lt)
    weights = CogvideoxTransformerWeights(cfg)
    weights.load_weights(W)
    infer = CogvideoxTransformerInfer(cfg)

    class Sched:
        image_rotary_emb = rotary

    infer.set_scheduler(Sched())
    with torch.no_grad():
        h_out, e_out = infer.infer(weights, hidden.clone(), enc.clone(), temb)
        blk = weights.blocks_weights[0]
        nh, ne, gate, enc_gate = infer.cogvideox_norm1(blk, hidden.clone(), enc.clone(), temb)
        ah, ae = infer.cogvideox_attention(blk, nh.clone(), ne.clone(), rotary)
    tensors = {"hidden_in": hidden, "enc_in": enc, "temb": temb, "cos": rotary[0], "sin": rotary[1], "hidden_out": h_out, "enc_out": e_out,
               "probe.norm1_hidden": nh, "probe.norm1_enc": ne, "probe.gate": gate, "probe.attn_hidden": ah, "probe.attn_enc": ae}
    meta = {"layers": str(layers), "heads": str(heads), "head_dim": str(hd), "ff": str(ff), "text_len": str(Lt), "grid": ",".join(map(str, grid)),
            "weights_seed": "42", "inputs_seed": "7", "rotary_seed": "3", "generator": "oracle/gen_golden.py:gen_cogvideox_fixture",
            "reference": "ModelTC/lightx2v@0591c35e"}
    save_file({k: v.contiguous() for k, v in tensors.items()}, os.path.join(GOLD, "cogvideox_2blocks.safetensors"), metadata=meta)
    print("cogvideox_2blocks hidden_out absmax", float(h_out.float().abs().max()), "bytes", os.path.getsize(os.path.join(GOLD, "cogvideox_2blocks.safetensors")))


def gen_hunyuan_prepost_fixture():
    """REAL HunyuanPreInfer methods (time_in / guidance_in / vector_in / img_in, pre_infer.py:68-80, 139-152) and the REAL HunyuanPostInfer
    (post_infer.py:11-33) with the reference's own weight classes (MM "Default" / "Default-Force-FP32", Conv3d "Default") on CPU at width
    1536, latent [1, 16, 3, 8, 12].  `infer_text_in` is left out: the reference's own call raises at this snapshot (see the oracle)."""
    from safetensors.torch import save_file

    import lightx2v.common.ops  # noqa: F401
    from lightx2v.common.ops import conv, mm  # noqa: F401
    from lightx2v.models.networks.hunyuan.infer.post_infer import HunyuanPostInfer
    from lightx2v.models.networks.hunyuan.infer.pre_infer import HunyuanPreInfer
    from lightx2v.models.networks.hunyuan.weights.post_weights import HunyuanPostWeights
    from lightx2v.models.networks.hunyuan.weights.pre_weights import HunyuanPreWeights

    from oracle import hunyuan_oracle as HO

    hidden = 1536
    cfg = Cfg(task="t2v", mm_config={}, cpu_offload=False)
    W = HO.synth_prepost_weights(hidden, seed=11)
    # the pre-weights tree also declares the token refiner's tensors: give them placeholders (never used by the methods run here)
    full = dict(W)
    pre_w = HunyuanPreWeights(cfg)

    def names(mod):
        for child in mod._modules.values():
            for a in ("weight_name", "bias_name"):
                n = getattr(child, a, None)
                if n is not None:
                    yield n

    for n in names(pre_w):
        if n not in full:
            full[n] = torch.zeros(8, 8, dtype=torch.bfloat16) if n.endswith("weight") else torch.zeros(8, dtype=torch.bfloat16)
    pre_w.load(full)
    post_w = HunyuanPostWeights(cfg)
    post_w.load(W)
    g = torch.Generator().manual_seed(5)
    latents = torch.randn(1, 16, 3, 8, 12, generator=g).to(torch.bfloat16)
    t = torch.tensor(713.0)
    guidance = torch.tensor([6.0], dtype=torch.bfloat16) * 1000.0
    text_states_2 = torch.randn(1, 768, generator=g).to(torch.bfloat16)
    pre = HunyuanPreInfer(cfg)
    time_out = pre.infer_time_in(pre_w, t)
    guid_out = pre.infer_guidance_in(pre_w, guidance)
    vec_out = pre.infer_vector_in(pre_w, text_states_2)
    img_out = pre.infer_img_in(pre_w, latents)
    vec = time_out + vec_out + guid_out
    img = torch.randn(3 * 4 * 6, hidden, generator=g).to(torch.bfloat16)

    class Sched:
        pass

    sched = Sched()
    sched.latents = latents
    post = HunyuanPostInfer(cfg)
    post.set_scheduler(sched)
    out = post.infer(post_w, img, vec)
    save_file({"latents": latents, "t": t.reshape(1), "guidance": guidance, "text_states_2": text_states_2, "time_out": time_out, "guidance_out": guid_out,
               "vector_out": vec_out, "img_out": img_out.contiguous(), "vec": vec, "img": img, "post_out": out.contiguous()},
              os.path.join(GOLD, "hunyuan_prepost.safetensors"),
              metadata={"hidden": str(hidden), "weights_seed": "11", "generator": "oracle/gen_golden.py:gen_hunyuan_prepost_fixture",
                        "reference": "ModelTC/lightx2v@0591c35e"})
    print("hunyuan_prepost post_out", tuple(out.shape), out.dtype, float(out.abs().max()))


def gen_vae_encode_fixture():
    """Real WanVAE_.encode (lightx2v/models/video_encoders/hf/wan/vae.py:684-711; Encoder3d :264-376) on CPU in fp32, seeded synthetic
    weights loaded through its own state_dict: video [3, 9, 32, 48] -> mu [16, 3, 4, 6]; pins oracle/vae_oracle.py:vae_encode."""
    from safetensors.torch import save_file

    from lightx2v.models.video_encoders.hf.wan.vae import WanVAE_

    from oracle import vae_oracle as V

    W = V.synth_vae_encoder_weights(seed=0)
    model = WanVAE_(dim=96, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[], temperal_downsample=[False, True, True], dropout=0.0).eval()
    res = model.load_state_dict(W, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.startswith("decoder.") or k.startswith("conv2.") for k in res.missing_keys), res.missing_keys
    g = torch.Generator().manual_seed(6)
    video = torch.rand(3, 9, 32, 48, generator=g) * 2 - 1
    scale = [torch.tensor(V.MEAN), 1.0 / torch.tensor(V.STD)]
    with torch.no_grad():
        mu = model.encode(video.unsqueeze(0), scale).float()[0]
    save_file({"video": video, "mu": mu.contiguous()}, os.path.join(GOLD, "wan_vae_encode_small.safetensors"),
              metadata={"weights_seed": "0", "generator": "oracle/gen_golden.py:gen_vae_encode_fixture", "reference": "ModelTC/lightx2v@0591c35e"})
    print("wan_vae_encode_small", tuple(mu.shape), "absmax", float(mu.abs().max()))


def gen_hunyuan_i2v_fixture():
    """Real HunyuanTransformerInfer with i2v token replacement (token_replace_vec, frist_frame_token_num; transformer_infer.py:100-103,
    192-209, 281-286, 316-328, 373-378): one double + one single block at width 3072, 96 image tokens of which the first 24 (the
    conditioning frame) follow the t = 0 embedding, 32 text tokens."""
    from safetensors.torch import save_file

    import lightx2v.common.ops  # noqa: F401
    from lightx2v.models.networks.hunyuan.infer.transformer_infer import HunyuanTransformerInfer
    from lightx2v.models.networks.hunyuan.weights.transformer_weights import HunyuanTransformerDoubleBlock, HunyuanTransformerSingleBlock

    from oracle import hunyuan_oracle as HO

    hidden, mlp, heads, first = 3072, 12288, 24, 24
    cfg = Cfg(cpu_offload=False, do_mm_calib=False, mm_config={}, attention_type="torch_sdpa", task="i2v")
    W = HO.synth_weights(1, 1, hidden, mlp, seed=42)
    img, txt, vec, cu, freqs = HO.synth_inputs(96, 32, 32, hidden, seed=7)
    trv = torch.randn(1, hidden, generator=torch.Generator().manual_seed(9)).to(torch.bfloat16)
    dbl, sgl = HunyuanTransformerDoubleBlock(0, cfg), HunyuanTransformerSingleBlock(0, cfg)
    dbl.load(W)
    sgl.load(W)
    infer = HunyuanTransformerInfer(cfg)
    cu_t = torch.tensor(cu, dtype=torch.int32)
    max_len = img.shape[0] + txt.shape[0]
    img1, txt1 = infer.infer_double_block(dbl, img.clone(), txt.clone(), vec, cu_t, max_len, freqs, trv, first)
    x = torch.cat((img1, txt1), 0)
    x2 = infer.infer_single_block(sgl, x, vec, txt.shape[0], cu_t, max_len, freqs, trv, first)
    save_file({"img": img, "txt": txt, "vec": vec, "token_replace_vec": trv, "cos": freqs[0], "sin": freqs[1], "img_after_double": img1.contiguous(),
               "txt_after_double": txt1.contiguous(), "x_after_single": x2.contiguous()},
              os.path.join(GOLD, "hunyuan_blocks_i2v_small.safetensors"),
              metadata={"hidden": str(hidden), "mlp": str(mlp), "heads": str(heads), "weights_seed": "42", "inputs_seed": "7", "txt_valid": "32",
                        "first_frame_tokens": str(first), "generator": "oracle/gen_golden.py:gen_hunyuan_i2v_fixture", "reference": "ModelTC/lightx2v@0591c35e"})
    print("hunyuan_blocks_i2v_small", float(x2.float().abs().max()), os.path.getsize(os.path.join(GOLD, "hunyuan_blocks_i2v_small.safetensors")))


if __name__ == "__main__":
    if os.environ.get("GOLDEN_ONLY", "") == "hunyuan_i2v":
        install_shims()
        os.makedirs(GOLD, exist_ok=True)
        gen_hunyuan_i2v_fixture()
        sys.exit(0)
    if os.environ.get("GOLDEN_ONLY", "") == "vae_encode":
        install_shims()
        os.makedirs(GOLD, exist_ok=True)
        gen_vae_encode_fixture()
        sys.exit(0)
    if os.environ.get("GOLDEN_ONLY", "") == "hunyuan_prepost":
        install_shims()
        os.makedirs(GOLD, exist_ok=True)
        gen_hunyuan_prepost_fixture()
        sys.exit(0)
    if os.environ.get("GOLDEN_ONLY", "") == "cogvideox":
        install_shims()
        os.makedirs(GOLD, exist_ok=True)
        gen_cogvideox_fixture()
        sys.exit(0)
    if os.environ.get("GOLDEN_ONLY", "") == "distill":
        install_shims()
        os.makedirs(GOLD, exist_ok=True)
        gen_distill_scheduler_fixture()
        sys.exit(0)
    if os.environ.get("GOLDEN_ONLY", "") == "prepost":
        install_shims()
        os.makedirs(GOLD, exist_ok=True)
        gen_prepost_fixture()
        sys.exit(0)
    if os.environ.get("GOLDEN_ONLY", "") == "causvid":
        install_shims()
        os.makedirs(GOLD, exist_ok=True)
        gen_causvid_fixture()
        gen_prepost_fixture()
        sys.exit(0)
    if os.environ.get("GOLDEN_ONLY", "") == "teacache":
        install_shims()
        os.makedirs(GOLD, exist_ok=True)
        gen_teacache_fixture()
        gen_causvid_fixture()
        gen_prepost_fixture()
        sys.exit(0)
    if os.environ.get("GOLDEN_ONLY", "") == "nvfp4":
        os.makedirs(GOLD, exist_ok=True)
        gen_nvfp4_fixture()
        gen_teacache_fixture()
        gen_causvid_fixture()
        gen_prepost_fixture()
        sys.exit(0)
    if os.environ.get("GOLDEN_ONLY", "") == "hunyuan_vae":
        install_shims()
        os.makedirs(GOLD, exist_ok=True)
        gen_hunyuan_vae_fixture()
        gen_nvfp4_fixture()
        gen_teacache_fixture()
        gen_causvid_fixture()
        gen_prepost_fixture()
        sys.exit(0)
    if os.environ.get("GOLDEN_ONLY", "") == "hunyuan":
        install_shims()
        os.makedirs(GOLD, exist_ok=True)
        gen_hunyuan_fixture()
    elif os.environ.get("GOLDEN_ONLY", "") == "vae":
        install_shims()
        os.makedirs(GOLD, exist_ok=True)
        gen_vae_fixture()
    elif os.environ.get("GOLDEN_ONLY", "") == "scheduler":
        install_shims()
        os.makedirs(GOLD, exist_ok=True)
        gen_scheduler_fixture()
    else:
        main()
        gen_scheduler_fixture()
        gen_distill_scheduler_fixture()
        gen_vae_fixture()
        gen_hunyuan_fixture()
        gen_hunyuan_vae_fixture()
        gen_nvfp4_fixture()
        gen_teacache_fixture()
        gen_causvid_fixture()
        gen_prepost_fixture()
        gen_cogvideox_fixture()
        gen_hunyuan_prepost_fixture()
        gen_vae_encode_fixture()
        gen_hunyuan_i2v_fixture()
