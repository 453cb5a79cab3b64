"""GPU parity of the pieces that bracket the block stack: WanPreInfer (patch-embed as a GEMM on unfolded
patches, time / text / CLIP MLPs), WanPostInfer (head LN + modulate + Linear + unpatchify) against the fixture of the REAL reference
classes, and WanModel.infer's cond / uncond / CFG combine against the oracle's formula on the model's own two passes."""
import os
import types

import pytest
import torch
from safetensors import safe_open

from oracle import wan_oracle as O

pytestmark = pytest.mark.gpu


def _load(path):
    with safe_open(path, framework="pt") as f:
        return {k: f.get_tensor(k) for k in f.keys()}, f.metadata()


def _bad_frac(got, ref, rtol=1e-2, atol=1e-2):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    return (err > atol + rtol * ref.abs()).float().mean().item(), err.max().item()


@pytest.mark.parametrize("task", ["t2v", "i2v"])
def test_pre_post_infer_vs_reference_fixture(golden_dir, task):
    from lightx2v_b200.host.wan_model import WanPostInfer, WanPreInfer

    T, meta = _load(os.path.join(golden_dir, "wan13b_prepost.safetensors"))
    dim = int(meta["dim"])
    g = lambda k: T[f"{task}.{k}"]       # noqa: E731
    cfg = dict(task=task, num_heads=12, dim=dim, freq_dim=256, text_len=512, out_dim=16, in_dim=36 if task == "i2v" else 16)
    W = {k: v.cuda() for k, v in O.synth_prepost_weights(dim, cfg["in_dim"], task, seed=int(meta["weights_seed"])).items()}
    sched = types.SimpleNamespace(latents=g("latents").cuda(), timesteps=g("timesteps").cuda(), step_index=int(meta["step_index"]), flag_df=False)
    inputs = {"text_encoder_output": {"context": g("context").cuda(), "context_null": g("context").cuda()}, "image_encoder_output": None}
    if task == "i2v":
        inputs["image_encoder_output"] = {"clip_encoder_out": g("clip_encoder_out").cuda(), "vae_encode_out": g("vae_encode_out").cuda()}
    pre, post = WanPreInfer(cfg), WanPostInfer(cfg)
    pre.set_scheduler(sched)
    post.set_scheduler(sched)
    embed, grid_sizes, (x, embed0, seq_lens, freqs, ctx) = pre.infer(W, inputs, True)
    torch.cuda.synchronize()
    assert grid_sizes[0].tolist() == g("grid").tolist()
    for name, got, ref in (("embed", embed, g("embed")), ("x", x, g("x")), ("embed0", embed0, g("embed0")), ("context", ctx, g("context_out"))):
        frac, mx = _bad_frac(got, ref)
        print(f"{task} pre_infer {name}: bad_frac={frac:.2e} max_abs_err={mx:.4f}")
        assert got.shape == ref.shape and frac < 2e-3, (name, frac, mx)
    noise = post.infer(W, g("x_blocks").cuda().clone(), g("embed").cuda(), grid_sizes)[0]
    frac, mx = _bad_frac(noise, g("noise_pred"))
    print(f"{task} post_infer: bad_frac={frac:.2e} max_abs_err={mx:.4f}")
    assert noise.dtype == torch.float32 and noise.shape == g("noise_pred").shape and frac < 2e-3


def test_model_cfg_combine_and_step():
    """WanModel.infer = cond pass, uncond pass, uncond + g (cond - uncond) in fp32 (wan/model.py:197-226): the combined prediction must equal
    the oracle's formula applied to the model's own two single passes, and one scheduler step must run end to end."""
    import bench as B
    from lightx2v_b200.host.wan_model import WanModel
    from lightx2v_b200.host.wan_scheduler import WanScheduler

    cfg = dict(dim=1536, num_heads=12, ffn_dim=8960, num_layers=2, target_shape=(16, 3, 16, 16), infer_steps=4, enable_cfg=True, sample_guide_scale=5.0,
               sample_shift=5.0, task="t2v", freq_dim=256, text_len=512, in_dim=16, out_dim=16, seed=42, mm_config={}, patch_size=(1, 2, 2))
    dev = torch.device("cuda")
    W = B.synth_weights(cfg, dev)
    gen = torch.Generator(device=dev).manual_seed(7)
    ctx = {"context": torch.randn(100, 4096, generator=gen, device=dev).to(torch.bfloat16), "context_null": torch.randn(512, 4096, generator=gen, device=dev).to(torch.bfloat16)}
    inputs = {"text_encoder_output": ctx, "image_encoder_output": None}
    model = WanModel.from_weight_dict(cfg, W)
    sched = WanScheduler(cfg, device=dev)
    sched.prepare()
    model.set_scheduler(sched)
    sched.step_pre(0)
    cond = model._forward(inputs, True).clone()
    uncond = model._forward(inputs, False).clone()
    model.infer(inputs)
    assert torch.equal(sched.noise_pred, O.cfg_combine(cond, uncond, 5.0))
    assert sched.noise_pred.dtype == torch.float32 and sched.noise_pred.shape == (16, 3, 16, 16)
    before = sched.latents.clone()
    sched.step_post()
    assert torch.isfinite(sched.latents).all() and not torch.equal(sched.latents.float(), before.float())


def _small_model(dev, scheduler_cls, steps=6):
    import bench as B
    from lightx2v_b200.host.wan_model import WanModel

    cfg = dict(dim=1536, num_heads=12, ffn_dim=8960, num_layers=2, target_shape=(16, 3, 16, 16), infer_steps=steps, enable_cfg=True, sample_guide_scale=5.0,
               sample_shift=5.0, task="t2v", freq_dim=256, text_len=512, in_dim=16, out_dim=16, seed=42, mm_config={}, patch_size=(1, 2, 2))
    W = B.synth_weights(cfg, dev)
    gen = torch.Generator(device=dev).manual_seed(7)
    ctx = {"context": torch.randn(100, 4096, generator=gen, device=dev).to(torch.bfloat16), "context_null": torch.randn(512, 4096, generator=gen, device=dev).to(torch.bfloat16)}
    inputs = {"text_encoder_output": ctx, "image_encoder_output": None}
    model = WanModel.from_weight_dict(cfg, W)
    sched = scheduler_cls(cfg, device=dev)
    sched.prepare()
    model.set_scheduler(sched)
    return model, sched, inputs


def test_device_scheduler_and_cuda_graph_step_are_bit_identical_to_the_eager_loop():
    """host/wan_graph.py: WanSchedulerDevice (coefficients from a device table, state in static buffers) driven eagerly, and the same
    steps replayed as CUDA graphs (one per step kind), must reproduce the eager WanScheduler loop bit for bit over a whole 6-step schedule
    (first step, second step, steady state, last step = all four graph kinds)."""
    from lightx2v_b200.host.wan_graph import GraphedDenoiser, WanSchedulerDevice
    from lightx2v_b200.host.wan_scheduler import WanScheduler

    dev = torch.device("cuda")
    steps = 6
    model, sched, inputs = _small_model(dev, WanScheduler, steps)
    want = []
    for i in range(steps):
        sched.step_pre(i)
        model.infer(inputs)
        sched.step_post()
        want.append(sched.latents.float().clone())
    # device scheduler, eager loop
    model_d, sched_d, inputs_d = _small_model(dev, WanSchedulerDevice, steps)
    assert sched_d.kinds == [(0, 1), (1, 2), (2, 2), (2, 2), (2, 2), (2, 1)]
    for i in range(steps):
        sched_d.load_step(i)
        sched_d.step_pre()
        model_d.infer(inputs_d)
        sched_d.step_post()
        assert torch.equal(sched_d.latents.float(), want[i]), (i, (sched_d.latents.float() - want[i]).abs().max())
    # CUDA graphs
    model_g, sched_g, inputs_g = _small_model(dev, WanSchedulerDevice, steps)
    den = GraphedDenoiser(model_g, sched_g, inputs_g)
    for i in range(steps):
        den.step(i)
        torch.cuda.synchronize()
        assert torch.equal(sched_g.latents.float(), want[i]), (i, (sched_g.latents.float() - want[i]).abs().max())
    assert len(den.graphs) == 4
