"""GPU parity of the fused DiT block path (host/wan_infer.py over libb200dit.so) against
(a) the committed fixtures produced by the REAL reference classes on CPU (oracle/gen_golden.py), and
(b) the oracle restatement executed on the same GPU with flash-attn + torch ops (the reference's own GPU path), and
(c) a float64 evaluation of the same blocks ("truth"): the CUDA path must be as close to it as the reference's bf16 path is.
Tolerance rtol = atol = 1e-2 (BASELINE.json north_star).  The residual stream has |x| ~ 1-4, where one bf16 ulp is 2^-7 .. 2^-6 absolute
(0.8 - 1.6 e-2): two correct bf16 pipelines with different fp32 summation orders land one ulp apart on a small fraction of elements, so
every test states the fraction it admits (set at ~3x what the code measures; the measured values are recorded to
gpurun_out/parity_numbers.json -> profiles/) and a hard cap on the largest error, and (c) shows the deviation is not a loss of accuracy."""
import os

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
    return ((got - ref).abs() > atol + rtol * ref.abs()).float().mean().item(), (got - ref).abs().max().item()


def _err_stats(got, truth):
    d = (got.double() - truth.double()).abs()
    return d.max().item(), d.pow(2).mean().sqrt().item()


def _build(cfg, W):
    from lightx2v_b200.host.wan_infer import WanTransformerInfer
    from lightx2v_b200.host.wan_weights import WanTransformerWeights

    weights = WanTransformerWeights(cfg)
    weights.load({k: v.cuda() for k, v in W.items()})
    return weights, WanTransformerInfer(cfg)


@pytest.mark.parametrize("name", ["wan13b_t2v_2blocks", "wan13b_i2v_1block"])
def test_blocks_vs_reference_fixture(golden_dir, name, record):
    T, meta = _load(os.path.join(golden_dir, name + ".safetensors"))
    dim, heads, ffn, L, task = int(meta["dim"]), int(meta["heads"]), int(meta["ffn"]), int(meta["layers"]), meta["task"]
    cfg = dict(task=task, num_layers=L, num_heads=heads, dim=dim, ffn_dim=ffn, mm_config={})
    W = O.synth_block_weights(L, dim, ffn, task=task, seed=int(meta["weights_seed"]))
    weights, infer = _build(cfg, W)
    grid = T["grid"].view(1, 3)
    freqs = O.wan_freqs_table(dim // heads)
    x = T["x_in"].cuda().clone()
    out = infer.infer(weights, grid, None, x, T["embed0"].cuda(), torch.tensor([x.shape[0]]), freqs, T["context"].cuda())
    torch.cuda.synchronize()
    frac, mx = _bad_frac(out, T["x_out"])
    print(f"{name}: bad_frac={frac:.3e} max_abs_err={mx:.4f}")
    record(bad_frac=frac, max_abs_err=mx)
    assert frac < 5e-4, (frac, mx)
    assert mx < 0.07


def test_per_phase_vs_reference_fixture(golden_dir, record):
    """Probe points of block 0: self-attention update, cross-attention update, FFN update, each from the fixture's own input."""
    T, meta = _load(os.path.join(golden_dir, "wan13b_t2v_2blocks.safetensors"))
    dim, heads, ffn = int(meta["dim"]), int(meta["heads"]), int(meta["ffn"])
    cfg = dict(task="t2v", num_layers=1, num_heads=heads, dim=dim, ffn_dim=ffn, mm_config={})
    W = O.synth_block_weights(2, dim, ffn, seed=int(meta["weights_seed"]))
    weights, infer = _build(cfg, W)
    blk = weights.blocks[0]
    freqs = O.wan_freqs_table(dim // heads)
    grid = T["grid"].view(1, 3)
    sh, sc, ga, csh, csc, cga = infer.infer_modulation(blk.compute_phases[0], T["embed0"].cuda())
    x = T["x_in"].cuda().clone()
    y = infer.infer_self_attn(blk.compute_phases[1], grid, x, None, freqs, sh, sc)          # reference-style return
    f, m = _bad_frac(y, T["probe.self_attn_y"])
    record(self_attn_y_bad_frac=f, self_attn_y_max=m)
    assert f < 5e-4, ("self_attn_y", f, m)
    x, attn_out = infer.infer_cross_attn(blk.compute_phases[2], x, T["context"].cuda(), y, ga)
    # fused path: x already holds x_after_cross + cross_attn_out
    ref = (T["probe.x_after_cross"].float() + T["probe.cross_attn_out"].float()).bfloat16()
    f, m = _bad_frac(x, ref)
    record(cross_bad_frac=f, cross_max=m)
    assert f < 5e-4, ("cross", f, m)
    yf = infer.infer_ffn(blk.compute_phases[3], x.clone(), None, csh, csc)                    # reference-style return
    f, m = _bad_frac(yf, T["probe.ffn_y"])
    record(ffn_y_bad_frac=f, ffn_y_max=m)
    assert f < 3e-3, ("ffn_y", f, m)


def test_block_vs_oracle_on_gpu_larger(record):
    """One 14B-width block (D 5120, 40 heads, F 13824) on 21x6x10 = 1260 tokens, against the oracle restatement run on the
    same GPU with flash-attn + torch ops (the reference's GPU path)."""
    dim, heads, ffn, grid = 5120, 40, 13824, (21, 6, 10)
    S = grid[0] * grid[1] * grid[2]
    W = O.synth_block_weights(1, dim, ffn, seed=1, device="cuda")
    x, embed0, context = O.synth_block_inputs(S, dim, seed=2, device="cuda")
    freqs = O.wan_freqs_table(128)
    ref = O.infer_blocks(W, 1, x.clone(), embed0, grid, freqs.cuda(), context, heads, attn="flash_attn2")
    cfg = dict(task="t2v", num_layers=1, num_heads=heads, dim=dim, ffn_dim=ffn, mm_config={})
    weights, infer = _build(cfg, W)
    out = infer.infer(weights, torch.tensor([grid]), None, x.clone(), embed0, None, freqs, context)
    torch.cuda.synchronize()
    f, m = _bad_frac(out, ref)
    print(f"14B-width block: bad_frac={f:.3e} max_abs_err={m:.4f}")
    record(bad_frac=f, max_abs_err=m)
    assert f < 1e-3 and m < 0.07, (f, m)


@pytest.mark.parametrize("dim,heads,ffn,layers,grid,task", [(1536, 12, 8960, 2, (3, 8, 10), "t2v"), (1536, 12, 8960, 1, (3, 8, 10), "i2v"),
                                                           (5120, 40, 13824, 1, (21, 6, 10), "t2v")])
def test_cuda_path_is_as_close_to_fp64_truth_as_the_reference_bf16_path(dim, heads, ffn, layers, grid, task, record):
    """The waiver behind the admitted bad fractions above: evaluate the same blocks in float64 (no intermediate rounding) and require
       err(cuda, truth) <= 1.1 x err(reference bf16 path, truth)      in RMS, and <= 1.5 x in max (a single-element statistic),
    where the reference bf16 path is the oracle restatement run on this GPU with flash-attn + torch ops."""
    S = grid[0] * grid[1] * grid[2]
    W = O.synth_block_weights(layers, dim, ffn, task=task, seed=21, device="cuda")
    x, embed0, context = O.synth_block_inputs(S, dim, task=task, seed=22, device="cuda")
    freqs = O.wan_freqs_table(dim // heads)
    truth = O.infer_blocks_exact(W, layers, x.clone(), embed0, grid, freqs, context, heads, task=task)
    ref = O.infer_blocks(W, layers, x.clone(), embed0, grid, freqs.cuda(), context, heads, task=task, attn="flash_attn2")
    cfg = dict(task=task, num_layers=layers, num_heads=heads, dim=dim, ffn_dim=ffn, mm_config={})
    weights, infer = _build(cfg, W)
    out = infer.infer(weights, torch.tensor([grid]), None, x.clone(), embed0, None, freqs, context)
    torch.cuda.synchronize()
    mx_o, rms_o = _err_stats(out, truth)
    mx_r, rms_r = _err_stats(ref, truth)
    print(f"{task} D={dim} L={layers}: cuda vs truth max {mx_o:.4f} rms {rms_o:.3e} | reference-bf16 vs truth max {mx_r:.4f} rms {rms_r:.3e}")
    record(cuda_max=mx_o, cuda_rms=rms_o, ref_bf16_max=mx_r, ref_bf16_rms=rms_r)
    assert rms_o <= 1.1 * rms_r, (rms_o, rms_r)
    assert mx_o <= 1.5 * mx_r, (mx_o, mx_r)


def test_cross_kv_cache_invalidates_on_new_context():
    dim, heads, ffn, grid = 1536, 12, 8960, (2, 4, 8)
    S = 64
    W = O.synth_block_weights(1, dim, ffn, seed=1, device="cuda")
    x, embed0, ctx1 = O.synth_block_inputs(S, dim, seed=2, device="cuda")
    _, _, ctx2 = O.synth_block_inputs(S, dim, seed=3, device="cuda")
    cfg = dict(task="t2v", num_layers=1, num_heads=heads, dim=dim, ffn_dim=ffn, mm_config={})
    weights, infer = _build(cfg, W)
    freqs = O.wan_freqs_table(128)
    g = torch.tensor([grid])
    o1 = infer.infer(weights, g, None, x.clone(), embed0, None, freqs, ctx1).clone()
    o2 = infer.infer(weights, g, None, x.clone(), embed0, None, freqs, ctx2).clone()
    o1b = infer.infer(weights, g, None, x.clone(), embed0, None, freqs, ctx1).clone()
    assert not torch.equal(o1, o2)
    assert torch.equal(o1, o1b)
    ctx1.mul_(0.5)                                  # in-place edit bumps the tensor version -> cache must refresh
    o1c = infer.infer(weights, g, None, x.clone(), embed0, None, freqs, ctx1)
    assert not torch.equal(o1, o1c)
    # a NEW tensor that the allocator places at the address of a freed one (same shape, version 0) must not alias the old entry:
    # the cache is keyed on the tensor object and keeps it alive, so this cannot happen by construction - checked here explicitly
    base = infer.infer(weights, g, None, x.clone(), embed0, None, freqs, ctx2).clone()
    tmp = ctx2.clone()
    o_tmp = infer.infer(weights, g, None, x.clone(), embed0, None, freqs, tmp).clone()
    addr = tmp.data_ptr()
    del tmp
    fresh = torch.empty_like(ctx2)
    fresh.copy_(ctx1 * 2.0)                           # different content, very likely the same address
    o_fresh = infer.infer(weights, g, None, x.clone(), embed0, None, freqs, fresh)
    assert torch.equal(o_tmp, base)
    assert not torch.equal(o_fresh, base), f"stale K/V reused (address reused: {fresh.data_ptr() == addr})"


@pytest.mark.parametrize("name", ["wan13b_t2v_2blocks", "wan13b_i2v_1block"])
def test_native_block_call_equals_per_op_schedule(golden_dir, name):
    """b200_wan_block_fwd (one C call per block, csrc/wan_block.cu) issues the same kernels in the same order as the per-op Python
    schedule: the two paths must agree bit for bit (t2v and the i2v two-softmax cross-attention)."""
    T, meta = _load(os.path.join(golden_dir, name + ".safetensors"))
    dim, heads, ffn, L, task = int(meta["dim"]), int(meta["heads"]), int(meta["ffn"]), int(meta["layers"]), meta["task"]
    W = O.synth_block_weights(L, dim, ffn, task=task, seed=int(meta["weights_seed"]))
    grid = T["grid"].view(1, 3)
    freqs = O.wan_freqs_table(dim // heads)
    outs = []
    for native in (True, False):
        cfg = dict(task=task, num_layers=L, num_heads=heads, dim=dim, ffn_dim=ffn, mm_config={}, b200_native_block=native)
        weights, infer = _build(cfg, W)
        assert infer.native_block == native
        x = T["x_in"].cuda().clone()
        outs.append(infer.infer(weights, grid, None, x, T["embed0"].cuda(), torch.tensor([x.shape[0]]), freqs, T["context"].cuda()).clone())
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    from lightx2v_b200 import lib
    assert lib.wan_block_workspace_bytes(100, 1536, 8960) == 100 * (1536 + 8960) * 2


def test_teacache_skips_blocks_and_reuses_the_cached_residual():
    """WanTransformerInferTeaCaching over the CUDA block stack: computed forwards equal the plain stack bit for bit and store x_out - x_in;
    skipped forwards return x + cached residual of the same (cond / uncond) branch.  Threshold 1e9 forces 'skip' between warm-up and cut-off."""
    from lightx2v_b200.host.wan_infer import WanTransformerInfer
    from lightx2v_b200.host.wan_teacache import WanTransformerInferTeaCaching

    dim, heads, ffn, L, grid = 1536, 12, 8960, 2, (2, 6, 8)
    S = grid[0] * grid[1] * grid[2]
    W = O.synth_block_weights(L, dim, ffn, seed=5)
    cfg = dict(task="t2v", num_layers=L, num_heads=heads, dim=dim, ffn_dim=ffn, mm_config={}, infer_steps=4, enable_cfg=True, teacache_thresh=1e9,
               coefficients=[[0, 0, 0, 1.0, 0.0], [0, 0, 0, 1.0, 0.0]], use_ret_steps=False)
    weights, plain = _build(cfg, W)
    tea = WanTransformerInferTeaCaching(cfg)

    class Sched:
        infer_steps = 4
        step_index = 0

    sched = Sched()
    tea.set_scheduler(sched)
    freqs = O.wan_freqs_table(dim // heads)
    g = torch.tensor([grid])
    pattern, residual = [], {True: None, False: None}
    for step in range(4):
        sched.step_index = step
        for cond in (True, False):
            x, embed0, context = O.synth_block_inputs(S, dim, seed=10 * step + int(cond), device="cuda")
            embed = embed0[:1].clone()
            assert tea.infer_conditional == cond
            out = tea.infer(weights, g, embed, x.clone(), embed0, None, freqs, context)
            computed = (sched.caching_records if cond else sched.caching_records_2)[step]
            pattern.append(computed)
            if computed:
                ref = plain.infer(weights, g, embed, x.clone(), embed0, None, freqs, context)
                assert torch.equal(out, ref)
                residual[cond] = ref - x
            else:
                assert torch.equal(out, x + residual[cond])
    assert pattern == [True, True, False, False, False, False, True, True]       # ret_steps = 2, cutoff = 2 * 4 - 2


def test_causvid_kv_cache_blocks_vs_reference_fixture(golden_dir):
    # (numbers of every chunk are printed; the fixture test above records the block-stack numbers)
    """WanTransformerInferCausVid on the CUDA kernels (K/V projected straight into the cache, RoPE at the chunk's frame offset, FMHA with
    sq != sk over the cache prefix) vs the fixture of the REAL reference class: three chunks, two blocks."""
    from lightx2v_b200.host.wan_causvid import WanTransformerInferCausVid
    from lightx2v_b200.host.wan_weights import WanTransformerWeights

    T, meta = _load(os.path.join(golden_dir, "wan13b_causvid_2blocks.safetensors"))
    dim, heads, ffn, L = int(meta["dim"]), int(meta["heads"]), int(meta["ffn"]), int(meta["layers"])
    chunks, ft = int(meta["chunks"]), int(meta["frame_tokens"])
    grid = torch.tensor([[int(v) for v in meta["grid"].split(",")]])
    cfg = dict(task="t2v", num_layers=L, num_heads=heads, dim=dim, ffn_dim=ffn, mm_config={}, num_frames=chunks, num_frame_per_block=1, frame_seq_length=ft,
               text_len=512)
    W = O.synth_block_weights(L, dim, ffn, seed=int(meta["weights_seed"]))
    weights = WanTransformerWeights(cfg)
    weights.load({k: v.cuda() for k, v in W.items()})
    infer = WanTransformerInferCausVid(cfg)
    infer._init_kv_cache(torch.bfloat16, "cuda")
    infer._init_crossattn_cache(torch.bfloat16, "cuda")
    freqs = O.wan_freqs_table(dim // heads)
    ctx = T["context"].cuda()
    for c in range(chunks):
        out = infer.infer(weights, grid, None, T[f"x_in.{c}"].cuda().clone(), T[f"embed0.{c}"].cuda(), torch.tensor([ft]), freqs, ctx, c * ft, (c + 1) * ft)
        torch.cuda.synchronize()
        frac, mx = _bad_frac(out, T[f"x_out.{c}"])
        print(f"causvid chunk {c}: bad_frac={frac:.3e} max_abs_err={mx:.4f}")
        assert frac < 1e-3 and mx < 0.07, (c, frac, mx)
    fk, mk = _bad_frac(infer.kv_cache[0]["k"].reshape(chunks * ft, dim), T["k_cache.0"])
    assert fk < 2e-3, (fk, mk)
