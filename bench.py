#!/usr/bin/env python
"""Headline benchmark: denoise-step latents/sec, Wan2.1-T2V-14B 720p x 81f (BASELINE.json configs[1]).

One "step" = scheduler.step_pre + model.infer (cond + uncond forwards, CFG combine; 2 x 40 DiT blocks over 75 600 tokens,
pre/post-infer included) + scheduler.step_post — the body of DefaultRunner.run's loop (lightx2v/models/runners/default_runner.py:97-114).
Synthetic latents / prompt embeddings / random-init weights of the named shapes (no datasets or checkpoints offline).

    python bench.py [--gpus N --steps K --warmup W]            our sm_100a path  (N > 1: torchrun, Ulysses over the token axis)
    python bench.py --impl reference [...]                     the reference's CPU torch path (oracle port) on the host cores,
                                                               bounded sample extrapolated by the FLOP model of SURVEY.md §8d
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "denoise-step latents/sec (Wan2.1-14B 720p x 81f)"
WORKLOADS = {
    # name: dims of the DiT + latent shape
    "wan2.1-t2v-14b-720p-81f": dict(dim=5120, num_heads=40, ffn_dim=13824, num_layers=40, target_shape=(16, 21, 90, 160), infer_steps=50,
                                    enable_cfg=True, sample_guide_scale=5.0, sample_shift=5.0),
    # BASELINE config 3: w8a8-fp8 linears (weights quantised per out-channel at load), 4-step distilled sampler, no CFG
    "wan2.1-t2v-14b-fp8-distill-720p-81f": dict(dim=5120, num_heads=40, ffn_dim=13824, num_layers=40, target_shape=(16, 21, 90, 160), infer_steps=4,
                                                enable_cfg=False, sample_guide_scale=1.0, sample_shift=5.0, fp8=True, distill=True),
    # north_star's w4a4-nvfp4 weight path on the same distilled sampler (the reference ships the kernels but no model wiring for it)
    "wan2.1-t2v-14b-nvfp4-distill-720p-81f": dict(dim=5120, num_heads=40, ffn_dim=13824, num_layers=40, target_shape=(16, 21, 90, 160), infer_steps=4,
                                                  enable_cfg=False, sample_guide_scale=1.0, sample_shift=5.0, nvfp4=True, distill=True),
    # BASELINE config 4: image-to-video (36 input channels = 16 noise + 4 mask + 16 VAE-encoded image, 257 CLIP tokens in a second cross-attention)
    "wan2.1-i2v-14b-720p-81f": dict(dim=5120, num_heads=40, ffn_dim=13824, num_layers=40, target_shape=(16, 21, 90, 160), infer_steps=50,
                                    enable_cfg=True, sample_guide_scale=5.0, sample_shift=5.0, task="i2v"),
    "wan2.1-t2v-1.3b-480p-17f": dict(dim=1536, num_heads=12, ffn_dim=8960, num_layers=30, target_shape=(16, 5, 60, 104), infer_steps=50,
                                     enable_cfg=True, sample_guide_scale=5.0, sample_shift=5.0),   # quick self-test of this script
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops_sustained"], d["hbm_gbs"], "MEASURED_PEAKS.json (sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def step_flops(cfg):
    from oracle.wan_oracle import block_flops  # FLOP model only; not on the measured path
    C, Fr, H, W = cfg["target_shape"]
    S = Fr * (H // 2) * (W // 2)
    per_fwd = cfg["num_layers"] * block_flops(S, cfg["dim"], cfg["ffn_dim"], 512, i2v=cfg.get("task") == "i2v")
    return S, per_fwd * (2 if cfg["enable_cfg"] else 1)


def synth_weights(cfg, device, seed=42):
    """Random-init weights with the checkpoint's key names and shapes (SURVEY.md §8d recipe), generated on the device."""
    g = torch.Generator(device=device).manual_seed(seed)
    D, F_, L = cfg["dim"], cfg["ffn_dim"], cfg["num_layers"]
    W = {}

    def rnd(*shape, scale=0.02):
        return (torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * scale).to(torch.bfloat16)

    def lin(name, n, k):
        W[name + ".weight"] = rnd(n, k)
        W[name + ".bias"] = rnd(n)

    i2v = cfg.get("task") == "i2v"
    W["patch_embedding.weight"] = rnd(D, 36 if i2v else 16, 1, 2, 2, scale=0.05)
    W["patch_embedding.bias"] = rnd(D)
    lin("text_embedding.0", D, 4096)
    lin("text_embedding.2", D, D)
    lin("time_embedding.0", D, 256)
    lin("time_embedding.2", D, D)
    lin("time_projection.1", 6 * D, D)
    lin("head.head", 64, D)
    W["head.modulation"] = rnd(1, 2, D, scale=0.1)
    for i in range(L):
        p = f"blocks.{i}."
        W[p + "modulation"] = rnd(1, 6, D, scale=0.1)
        for nm in ("q", "k", "v", "o"):
            lin(p + "self_attn." + nm, D, D)
            lin(p + "cross_attn." + nm, D, D)
        for nm in ("self_attn.norm_q", "self_attn.norm_k", "cross_attn.norm_q", "cross_attn.norm_k", "norm3"):
            W[p + nm + ".weight"] = (1.0 + rnd(D, scale=0.05).float()).to(torch.bfloat16)
        W[p + "norm3.bias"] = rnd(D)
        lin(p + "ffn.0", F_, D)
        lin(p + "ffn.2", D, F_)
        if i2v:
            lin(p + "cross_attn.k_img", D, D)
            lin(p + "cross_attn.v_img", D, D)
            W[p + "cross_attn.norm_k_img.weight"] = (1.0 + rnd(D, scale=0.05).float()).to(torch.bfloat16)
    if i2v:                                                       # img_emb.proj: LN(1280) -> Linear(1280, 1280) -> GELU -> Linear(1280, D) -> LN(D)
        W["img_emb.proj.0.weight"], W["img_emb.proj.0.bias"] = (1.0 + rnd(1280, scale=0.05).float()).to(torch.bfloat16), rnd(1280)
        lin("img_emb.proj.1", 1280, 1280)
        lin("img_emb.proj.3", D, 1280)
        W["img_emb.proj.4.weight"], W["img_emb.proj.4.bias"] = (1.0 + rnd(D, scale=0.05).float()).to(torch.bfloat16), rnd(D)
    return W


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def run_ours(args):
    from lightx2v_b200 import lib
    from lightx2v_b200.host import ulysses
    from lightx2v_b200.host.wan_model import WanModel
    from lightx2v_b200.host.wan_scheduler import WanScheduler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lib.load()

    cfg = dict(WORKLOADS[args.workload])
    cfg.update(task=cfg.get("task", "t2v"), freq_dim=256, text_len=512, in_dim=36 if cfg.get("task") == "i2v" else 16, out_dim=16, seed=42, mm_config={},
               patch_size=(1, 2, 2))
    if cfg.get("fp8"):
        from lightx2v_b200.host.ops import FP8_MM_KEY
        cfg["mm_config"] = {"mm_type": FP8_MM_KEY, "weight_auto_quant": True}
    if cfg.get("nvfp4"):
        from lightx2v_b200.host.ops import NVFP4_MM_KEY
        cfg["mm_config"] = {"mm_type": NVFP4_MM_KEY}
    if cfg.get("distill"):
        cfg["denoising_step_list"] = [1000, 750, 500, 250]
    # per-op launch schedule from Python, so that every kernel launch is counted and the FMHA launches carry CUDA events for the roofline;
    # the library's default (one native b200_wan_block_fwd call per block) issues the same kernels in the same order (tests/test_gpu_block.py)
    cfg["b200_native_block"] = False
    S, flops_step = step_flops(cfg)
    W = synth_weights(cfg, dev)
    model = WanModel(cfg, W)
    if cfg.get("distill"):
        from lightx2v_b200.host.wan_scheduler import WanStepDistillScheduler
        sched = WanStepDistillScheduler(cfg, device=dev)
    else:
        sched = WanScheduler(cfg, device=dev)
    sched.prepare()
    model.set_scheduler(sched)
    g = torch.Generator(device=dev).manual_seed(7)
    ctx = {"context": torch.randn(512, 4096, generator=g, device=dev).to(torch.bfloat16),
           "context_null": torch.randn(512, 4096, generator=g, device=dev).to(torch.bfloat16)}
    inputs = {"text_encoder_output": ctx, "image_encoder_output": None}
    if cfg["task"] == "i2v":      # synthetic encoder outputs of the published shapes: CLIP ViT-H tokens, VAE-encoded image + mask
        ts = cfg["target_shape"]
        inputs["image_encoder_output"] = {"clip_encoder_out": torch.randn(257, 1280, generator=g, device=dev).to(torch.bfloat16),
                                          "vae_encode_out": torch.randn(20, ts[1], ts[2], ts[3], generator=g, device=dev).to(torch.bfloat16)}

    # launch counter + per-launch CUDA events for the dominant kernel (self-attention FMHA)
    counters = {"launches": 0}
    fmha_events = []
    timing = {"on": False}
    native = {n: getattr(lib, n) for n in ("gemm_bf16", "ln_modulate", "rms_rope_", "fmha", "rms_rope_scatter", "fmha_scatter", "gemm_fp8",
                                           "quant_fp8_per_token", "ln_modulate_fp8", "gemm_nvfp4", "quant_nvfp4", "nvfp4_act_scale")}

    def counted(name):
        fn = native[name]

        def wrapper(*a, **k):
            counters["launches"] += 1
            if name in ("fmha", "fmha_scatter") and timing["on"] and a[0].shape[0] == a[1].shape[0]:
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                r = fn(*a, **k)
                e.record()
                fmha_events.append((s, e, a[0].shape[0], a[1].shape[0], a[0].shape[1]))
                return r
            return fn(*a, **k)

        return wrapper

    for n in native:
        setattr(lib, n, counted(n))
    if world > 1:
        # after wrapping, so the sharded launches are counted and timed too
        sp_mode = args.sp
        if args.parallel == "cfg" and cfg["enable_cfg"] and world % 2 == 0:
            sp_mode = ulysses.parallelize_wan_cfg(model, S, lib.fmha, sp=args.sp)
        elif sp_mode == "fused":
            try:
                ulysses.parallelize_wan_fused(model, S)
            except Exception as ex:  # symmetric memory unavailable on this box: keep the run alive on the NCCL exchange and say so
                sp_mode = f"nccl (peer-memory setup failed: {str(ex)[:120]})"
                ulysses.parallelize_wan(model, S, lib.fmha)
        else:
            ulysses.parallelize_wan(model, S, lib.fmha)
    else:
        sp_mode = "none"

    def one_step(i):
        i = i % max(1, sched.infer_steps - 1)
        if i == 0 and not cfg.get("distill"):
            sched.set_timesteps(sched.infer_steps, shift=sched.sample_shift)   # fresh multistep history when the 50-step grid wraps
        sched.step_pre(i)
        model.infer(inputs)
        sched.step_post()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    step_idx = 0
    for _ in range(args.warmup):
        one_step(step_idx)
        step_idx += 1
    barrier()

    # ---------------- timed region 1: inputs resident in HBM
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    counters["launches"] = 0
    timing["on"] = True
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.nvtx.range_push("timed")          # ncu --nvtx --nvtx-include "timed/" captures exactly this region
    t0.record()
    for _ in range(args.steps):
        one_step(step_idx)
        step_idx += 1
    t1.record()
    torch.cuda.nvtx.range_pop()
    barrier()
    timing["on"] = False
    ms_resident = t0.elapsed_time(t1) / args.steps
    launches = counters["launches"]
    clk = clocks.stop() if rank == 0 else None

    # ---------------- timed region 2: end to end through the public API with HOST buffers (H2D inputs, D2H result per step)
    h_lat = torch.empty(cfg["target_shape"], dtype=torch.float32).pin_memory()
    h_lat.copy_(sched.latents.float().cpu())
    h_ctx = {k: v.cpu().pin_memory() for k, v in ctx.items()}
    h_out = torch.empty(cfg["target_shape"], dtype=torch.float32).pin_memory()
    h2d = h_lat.numel() * 4 + sum(v.numel() * 2 for v in h_ctx.values())
    d2h = h_out.numel() * 4
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sched.latents = h_lat.to(dev, non_blocking=True)
        inputs["text_encoder_output"] = {k: v.to(dev, non_blocking=True) for k, v in h_ctx.items()}
        one_step(step_idx)
        step_idx += 1
        h_out.copy_(sched.latents.float(), non_blocking=True)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps

    # max over ranks
    if world > 1:
        t = torch.tensor([ms_resident, ms_e2e], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_resident, ms_e2e = t.tolist()

    if rank == 0:
        peak_tf, peak_gbs, peak_src = peaks()
        fm = [(s.elapsed_time(e), sq, sk, h) for s, e, sq, sk, h in fmha_events]
        roof = None
        if fm:
            avg_ms = sum(x[0] for x in fm) / len(fm)
            _, sq, sk, h = fm[0]
            fl = 4.0 * sq * sk * h * 128
            ach = fl / (avg_ms * 1e-3) / 1e12
            # DRAM bytes per launch from the committed `ncu --set full` capture of this kernel at this shape
            # (profiles/r01_fmha_v4_h40_ncu_summary.txt: dram read 2.70 GB + write 1.09 GB; algorithmic q+k+v+o = 3.10 GB).
            traffic = 3.787e9 if (sq, sk, h) == (75600, 75600, 40) else None
            roof = {"kernel": "fmha_fwd_d128_kernel (self-attention)", "bound": "tensor", "achieved": round(ach, 1), "peak": peak_tf,
                    "unit": "TFLOP/s", "frac": round(ach / peak_tf, 4), "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write)",
                    "launch_ms": round(avg_ms, 3),
                    "launches_timed": len(fm), "algorithmic_flop_per_launch": fl, "peak_source": peak_src}
        out = {
            "metric": METRIC, "value": round(1000.0 / ms_resident, 5), "unit": "latents/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_resident, 2), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "fp8-e4m3 linears, bf16 attention" if cfg.get("fp8") else ("nvfp4 (e2m1 + ue4m3/16) linears, bf16 attention" if cfg.get("nvfp4") else "bf16"),
            "data": "synthetic latents/prompt embeddings, random-init weights of the named shapes",
            "config": {"workload": args.workload, "tokens": S, "forwards_per_step": 2 if cfg["enable_cfg"] else 1, "blocks": cfg["num_layers"],
                       "parallelism": (sp_mode if sp_mode.startswith("cfg2") else f"ulysses{world}") if world > 1 else "single", "sp_exchange": sp_mode, "block_schedule": "per-op launches (instrumented)", "l2": "activations (774 MB/tensor) and weights (28 GB) exceed the 126 MB L2",
                       "scheduler": "step-distill 4-step (x0 re-noising)" if cfg.get("distill") else "UniPC order 2 (flow), 50-step sigma grid"},
            "achieved_tflops": round(flops_step / (ms_resident * 1e-3) / 1e12, 1),
            "model_tflop_per_step": round(flops_step / 1e12, 1),
            "e2e": {"value": round(1000.0 / ms_e2e, 5), "unit": "latents/s", "ms_per_step": round(ms_e2e, 2), "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": roof,
        }
        if args.vae and world == 1:
            del model, W
            sched.latents = None
            torch.cuda.empty_cache()
            out["vae_decode"] = vae_decode_bench(cfg, dev, with_reference=args.gpu_reference)
        if args.cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_reference_sample(cfg, flops_step, budget_s=args.cpu_budget)
        if args.gpu_reference and world == 1:
            out["gpu_reference"] = gpu_reference_sample(cfg, S, dev)
        print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


def run_hunyuan_blocks(args):
    """BASELINE config 5 (block stack only): HunyuanVideo 13B bf16, 720p x 129f = 118 800 image tokens + 256 text tokens (77 valid),
    20 double-stream + 40 single-stream blocks, one forward per step (embedded guidance, no CFG).  Pre/post-infer and the Hunyuan VAE
    are not part of this measurement (SURVEY.md 8a A16 only)."""
    from lightx2v_b200 import lib
    from lightx2v_b200.host.hunyuan_infer import HunyuanTransformerInfer, HunyuanTransformerWeights

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    lib.load()
    D, M, H, ND, NS = 3072, 12288, 24, 20, 40
    Li, Lt, valid = 33 * 45 * 80, 256, 77
    g = torch.Generator(device=dev).manual_seed(42)

    def rnd(*shape, scale=0.02):
        return (torch.randn(*shape, generator=g, device=dev, dtype=torch.float32) * scale).to(torch.bfloat16)

    W = {}
    def lin(name, n, k, scale=0.02):
        W[name + ".weight"] = rnd(n, k, scale=scale)
        W[name + ".bias"] = rnd(n)
    for i in range(ND):
        p = f"double_blocks.{i}."
        for s in ("img", "txt"):
            lin(p + s + "_mod.linear", 6 * D, D, 0.01); lin(p + s + "_attn_qkv", 3 * D, D); lin(p + s + "_attn_proj", D, D)
            lin(p + s + "_mlp.fc1", M, D); lin(p + s + "_mlp.fc2", D, M)
            W[p + s + "_attn_q_norm.weight"] = (1 + rnd(128, scale=0.05).float()).to(torch.bfloat16)
            W[p + s + "_attn_k_norm.weight"] = (1 + rnd(128, scale=0.05).float()).to(torch.bfloat16)
    for i in range(NS):
        p = f"single_blocks.{i}."
        lin(p + "linear1", 3 * D + M, D); lin(p + "linear2", D, D + M); lin(p + "modulation.linear", 3 * D, D, 0.01)
        W[p + "q_norm.weight"] = (1 + rnd(128, scale=0.05).float()).to(torch.bfloat16)
        W[p + "k_norm.weight"] = (1 + rnd(128, scale=0.05).float()).to(torch.bfloat16)
    cfg = dict(task="t2v", mm_config={}, double_blocks_num=ND, single_blocks_num=NS)
    weights = HunyuanTransformerWeights(cfg)
    weights.load(W)
    infer = HunyuanTransformerInfer(cfg)
    img0, txt0, vec = rnd(Li, D, scale=1.0), rnd(Lt, D, scale=1.0), rnd(1, D, scale=1.0)
    ang = torch.rand(Li, 64, generator=g, device=dev) * 6.28
    freqs = (ang.cos().repeat_interleave(2, 1).to(torch.bfloat16), ang.sin().repeat_interleave(2, 1).to(torch.bfloat16))
    cu = [0, Li + valid, Li + Lt]

    def step():
        infer.infer(weights, img0.clone(), txt0.clone(), vec, cu, Li + Lt, freqs)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / args.steps
    L = Li + Lt
    flops = 60 * (4.0 * L * L * D + 24.0 * L * D * D)
    print(json.dumps({"metric": "denoise-step latents/sec (HunyuanVideo 13B 720p x 129f, DiT block stack only)", "value": round(1000.0 / ms, 5),
                      "unit": "latents/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 1), "higher_is_better": True,
                      "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": args.workload, "img_tokens": Li, "txt_tokens": Lt, "txt_valid": valid, "blocks": "20 double + 40 single"},
                      "achieved_tflops": round(flops / (ms * 1e-3) / 1e12, 1), "model_tflop_per_step": round(flops / 1e12, 1),
                      "vae_decode": hunyuan_vae_decode_bench(dev) if args.vae else None}))


def hunyuan_vae_decode_bench(dev):
    """HunyuanVideo VAE decode of the 720p x 129f latent [1,16,33,90,160] with the reference's tiling (3 temporal x 28 spatial tiles),
    through HunyuanVAEB200.decode (device->host copy of the fp32 video included in `wall_s`), next to the torch restatement of the
    reference's fp16 cuDNN path on ONE full tile [16,17,32,32] of the same GPU (bounded sample; diffusers is absent on the box, so the
    reference classes themselves cannot be imported there)."""
    import time

    from oracle import hunyuan_vae_oracle as HV
    from lightx2v_b200.host.hunyuan_vae import HunyuanVAEB200

    cfg = dict(HV.HUNYUAN_VAE_CFG)
    W = {k: v.to(dev) for k, v in HV.synth_vae_weights(cfg, seed=5).items()}
    vae = HunyuanVAEB200(W, device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    lat = torch.randn(1, 16, 33, 90, 160, generator=g, device=dev)
    vae.decode_device(lat[:, :, :5, :32, :32])                                   # warm-up (kernel attribute setup, allocator)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    s.record()
    img = vae.decode_device(lat)
    e.record()
    out = img.cpu().float()
    wall = time.time() - t0
    ms = s.elapsed_time(e)
    mpix = out.shape[2] * out.shape[3] * out.shape[4] / 1e6
    res = {"unit": "MPix/s", "output": list(out.shape[1:]), "mpix": round(mpix, 2), "value": round(mpix / (ms * 1e-3), 1), "ms": round(ms, 1),
           "wall_s_with_d2h": round(wall, 2), "tiles": "3 temporal x 4 x 7 spatial (25 % overlap, linear blends)",
           "dtype": "bf16 activations, fp32 accumulate, fp32/fp64 GroupNorm statistics", "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
    tile = lat[:, :, :17, :32, :32]
    s.record()
    vae.decoder.decode_tile(tile[0])
    e.record()
    torch.cuda.synchronize()
    ms_tile = s.elapsed_time(e)
    Wh = {k: v.half() for k, v in W.items()}
    with torch.no_grad():
        zt = (tile / cfg["scaling_factor"]).half()
        HV.tile_decode(Wh, zt[:, :, :3, :8, :8], cfg)                            # cuDNN warm-up
        torch.cuda.synchronize()
        s.record()
        HV.tile_decode(Wh, zt, cfg)
        e.record()
        torch.cuda.synchronize()
    ms_ref = s.elapsed_time(e)
    res["gpu_reference"] = {"sample": "one full tile [16,17,32,32] -> [3,65,256,256], torch fp16 (cuDNN) restatement of the reference decoder",
                            "ms_tile_reference": round(ms_ref, 1), "ms_tile_ours": round(ms_tile, 1), "speedup": round(ms_ref / ms_tile, 2),
                            "value": round(mpix / (ms * 1e-3) * ms_tile / ms_ref, 1), "unit": "MPix/s (extrapolated by the tile ratio)"}
    return res


def cpu_reference_sample(cfg, flops_step, budget_s=20.0):
    """The reference's CPU torch path (oracle port, pinned to the real reference by tests/golden) on the host cores:
    ONE block of the workload's width on a bounded token count, extrapolated to the full step by the FLOP model."""
    from oracle import wan_oracle as O
    D, F_, H = cfg["dim"], cfg["ffn_dim"], cfg["num_heads"]
    grid = (4, 16, 16)                                      # 1024 tokens
    S = grid[0] * grid[1] * grid[2]
    W = O.synth_block_weights(1, D, F_, seed=1)
    x, embed0, context = O.synth_block_inputs(S, D, seed=2)
    freqs = O.wan_freqs_table(128)
    # use the thread count that serves the reference best on this host (all cores is not always the fastest for torch CPU ops)
    ncpu = os.cpu_count() or 1
    best, threads = None, ncpu
    for cand in sorted({ncpu, min(ncpu, 32)}, reverse=True):
        torch.set_num_threads(cand)
        O.infer_blocks(W, 1, x.clone(), embed0, grid, freqs, context, H)
        t0 = time.time()
        O.infer_blocks(W, 1, x.clone(), embed0, grid, freqs, context, H)
        dt = time.time() - t0
        if best is None or dt < best:
            best, threads = dt, cand
    torch.set_num_threads(threads)
    O.infer_blocks(W, 1, x.clone(), embed0, grid, freqs, context, H)          # warm-up
    n, t0 = 0, time.time()
    while True:
        O.infer_blocks(W, 1, x.clone(), embed0, grid, freqs, context, H)
        n += 1
        if time.time() - t0 > budget_s or n >= 5:
            break
    sec = (time.time() - t0) / n
    fl = O.block_flops(S, D, F_, 512)
    tflops = fl / sec / 1e12
    est_step_s = flops_step / (tflops * 1e12)
    return {"value": 1.0 / est_step_s, "unit": "latents/s", "cores": threads, "kind": "port",
            "sample": f"1 DiT block (D={D}, F={F_}) at {S} tokens, {n} runs, {sec:.2f} s/block = {tflops:.3f} TFLOP/s on {threads} threads; "
                      f"extrapolated to the {flops_step / 1e12:.0f} TFLOP step by the FLOP model (SURVEY.md 8d)"}


def vae_decode_bench(cfg, dev, with_reference=True):
    """Second half of BASELINE.json's metric: Wan VAE decode MPix/s on the workload's latent ([16, 21, 90, 160] -> 81 x 720 x 1280)."""
    from lightx2v_b200.host.wan_vae import WanVAEDecoderB200
    from oracle import vae_oracle as V           # synthetic weights + (optional) the reference decode loop as GPU baseline
    C, Fr, Hh, Ww = cfg["target_shape"]
    Wd = V.synth_vae_weights(0)
    dec = WanVAEDecoderB200(Wd, device=dev)
    g = torch.Generator(device=dev).manual_seed(3)
    zs = torch.randn(C, Fr, Hh, Ww, generator=g, device=dev)
    mpix = (1 + 4 * (Fr - 1)) * Hh * 8 * Ww * 8 / 1e6
    res = {"unit": "MPix/s", "output": [3, 1 + 4 * (Fr - 1), Hh * 8, Ww * 8], "mpix": round(mpix, 2), "dtype": "bf16 activations, fp32 accumulate"}
    try:
        img = dec.decode(zs)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        n = 2
        for _ in range(n):
            img = dec.decode(zs)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / n
        res.update(value=round(mpix / (ms * 1e-3), 1), ms=round(ms, 1), effective_tflops=round(639.3 / (ms * 1e-3), 1),   # the reference algorithm's 639.3 TFLOP / time
                   peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2**30, 1))
        del img
    except Exception as ex:  # noqa
        res["error"] = str(ex)[:300]
    if with_reference:
        try:
            del dec
            torch.cuda.empty_cache()
            Wg = {k: v.to(dev) for k, v in Wd.items()}
            zs_small = zs[:, :6]                                   # the reference loop is per latent frame: 6 frames -> 21 video frames
            V.vae_decode(Wg, zs_small[:, :2])
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            V.vae_decode(Wg, zs_small)
            e.record()
            torch.cuda.synchronize()
            frames = 1 + 4 * (zs_small.shape[1] - 1)
            mp = frames * Hh * 8 * Ww * 8 / 1e6
            res["gpu_reference"] = {"value": round(mp / (s.elapsed_time(e) * 1e-3), 1), "unit": "MPix/s",
                                    "sample": f"reference per-frame decode loop (fp32, cuDNN TF32), {zs_small.shape[1]} latent frames -> {frames} frames"}
        except Exception as ex:  # noqa
            res["gpu_reference"] = {"unavailable": str(ex)[:200]}
    return res


def gpu_reference_sample(cfg, S, dev):
    """The reference's own GPU path (flash-attn 2 + torch ops, oracle restatement executed on the GPU): one block at the
    full token count, extrapolated x blocks x forwards.  Extra information beside the contract's CPU reference arm."""
    from oracle import wan_oracle as O
    D, F_, H, L = cfg["dim"], cfg["ffn_dim"], cfg["num_heads"], cfg["num_layers"]
    C, Fr, Hh, Ww = cfg["target_shape"]
    grid = (Fr, Hh // 2, Ww // 2)
    try:
        W = O.synth_block_weights(1, D, F_, seed=1, device=dev)
        x, embed0, context = O.synth_block_inputs(S, D, seed=2, device=dev)
        freqs = O.wan_freqs_table(128).to(dev)
        for _ in range(1):
            O.infer_blocks(W, 1, x.clone(), embed0, grid, freqs, context, H, attn="flash_attn2")
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        n = 2
        for _ in range(n):
            O.infer_blocks(W, 1, x.clone(), embed0, grid, freqs, context, H, attn="flash_attn2")
        e.record()
        torch.cuda.synchronize()
        ms_block = s.elapsed_time(e) / n
        fw = 2 if cfg["enable_cfg"] else 1
        return {"value": 1000.0 / (ms_block * L * fw), "unit": "latents/s", "ms_per_block": round(ms_block, 2),
                "sample": f"1 block at {S} tokens with flash_attn_varlen_func + torch.addmm/layer_norm (reference op order), x{L} blocks x{fw} forwards"}
    except Exception as ex:  # noqa
        return {"unavailable": str(ex)[:200]}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cfg = dict(WORKLOADS[args.workload])
    S, flops_step = step_flops(cfg)
    vals = []
    for _ in range(args.warmup + args.steps):
        vals.append(cpu_reference_sample(cfg, flops_step, budget_s=max(2.0, args.cpu_budget / 2)))
    timed = vals[args.warmup:] or vals
    v = sum(x["value"] for x in timed) / len(timed)
    base = dict(timed[-1])
    base["value"] = v
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "latents/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
           "data": "synthetic", "config": {"workload": args.workload, "tokens": S},
           "cpu_baseline": base, "e2e": {"value": v, "unit": "latents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="wan2.1-t2v-14b-720p-81f", choices=list(WORKLOADS) + ["hunyuan-13b-720p-129f-blocks"])
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-gpu-reference", dest="gpu_reference", action="store_false")
    ap.add_argument("--no-vae", dest="vae", action="store_false")
    ap.add_argument("--sp", default="fused", choices=["fused", "nccl"], help="Ulysses exchange: peer-memory kernels (default) or NCCL all-to-all")
    ap.add_argument("--parallel", default="ulysses", choices=["ulysses", "cfg"],
                    help="N > 1: Ulysses over all ranks, or CFG-parallel (cond / uncond on rank halves) x Ulysses inside each half")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    args = ap.parse_args()
    if args.workload == "hunyuan-13b-720p-129f-blocks":
        run_hunyuan_blocks(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
