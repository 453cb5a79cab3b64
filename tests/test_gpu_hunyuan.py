"""GPU parity of the HunyuanVideo block path (host/hunyuan_infer.py over libb200dit.so): per-head RMS + bf16-chain RoPE kernel vs the
oracle ops, one double + one single block vs the fixture produced by the REAL HunyuanTransformerInfer, and a two-segment varlen case
(valid + padded text tokens) vs the oracle run on the GPU with flash-attn.  Tolerance rtol = atol = 1e-2 with the one-ulp allowance
described in tests/test_gpu_block.py."""
import os

import pytest
import torch
from safetensors import safe_open

from oracle import hunyuan_oracle as HO

pytestmark = pytest.mark.gpu


def _bad(got, ref, rtol=1e-2, atol=1e-2):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    return (err > atol + rtol * ref.abs()).float().mean().item(), err.max().item()


def _build(W, n_double, n_single):
    from lightx2v_b200.host.hunyuan_infer import HunyuanTransformerInfer, HunyuanTransformerWeights

    cfg = dict(task="t2v", mm_config={}, double_blocks_num=n_double, single_blocks_num=n_single)
    weights = HunyuanTransformerWeights(cfg)
    weights.load({k: v.cuda() for k, v in W.items()})
    return weights, HunyuanTransformerInfer(cfg)


def test_rms_rope_heads_kernel():
    from lightx2v_b200 import lib
    from lightx2v_b200.host.hunyuan_infer import rope_cos_sin_pairs

    g = torch.Generator(device="cuda").manual_seed(1)
    L, H = 300, 24
    qkv = (torch.randn(L, 3, H, 128, generator=g, device="cuda") * 2).to(torch.bfloat16)
    wq = (1 + 0.1 * torch.randn(128, generator=g, device="cuda")).to(torch.bfloat16)
    wk = (1 + 0.1 * torch.randn(128, generator=g, device="cuda")).to(torch.bfloat16)
    ang = torch.rand(200, 64, generator=g, device="cuda") * 6.28
    cos, sin = ang.cos().repeat_interleave(2, 1).to(torch.bfloat16), ang.sin().repeat_interleave(2, 1).to(torch.bfloat16)
    q, k = qkv[:, 0].clone(), qkv[:, 1].clone()
    rq, rk = HO.rms_norm(q, wq), HO.rms_norm(k, wk)
    rq2, rk2 = HO.apply_rotary_emb(rq[:200], rk[:200], (cos, sin))
    rq, rk = torch.cat((rq2, rq[200:])), torch.cat((rk2, rk[200:]))
    lib.rms_rope_heads_(qkv[:, 0], wq, qkv[:, 1], wk, cos_sin=rope_cos_sin_pairs((cos, sin)), rope_rows=200)
    for got, ref in ((qkv[:, 0], rq), (qkv[:, 1], rk)):
        f, m = _bad(got, ref)
        assert f < 2e-3, (f, m)
    assert torch.equal(qkv[:, 2], qkv[:, 2])


def test_blocks_vs_reference_fixture(golden_dir):
    with safe_open(os.path.join(golden_dir, "hunyuan_blocks_small.safetensors"), framework="pt") as f:
        T = {k: f.get_tensor(k) for k in f.keys()}
        meta = f.metadata()
    hidden, mlp = int(meta["hidden"]), int(meta["mlp"])
    W = HO.synth_weights(1, 1, hidden, mlp, seed=int(meta["weights_seed"]))
    weights, infer = _build(W, 1, 1)
    Li, Lt = T["img"].shape[0], T["txt"].shape[0]
    cu = [0, Li + int(meta["txt_valid"]), Li + Lt]
    freqs = (T["cos"].cuda(), T["sin"].cuda())
    img, txt = infer.infer_double_block(weights.double_blocks[0], T["img"].cuda().clone(), T["txt"].cuda().clone(), T["vec"].cuda(), cu, Li + Lt, freqs)
    f1, m1 = _bad(img, T["img_after_double"])
    f2, m2 = _bad(txt, T["txt_after_double"])
    print(f"hunyuan double block: img bad {f1:.2e} max {m1:.4f}; txt bad {f2:.2e} max {m2:.4f}")
    assert f1 < 2e-3 and f2 < 2e-3 and max(m1, m2) < 0.13
    x = infer.infer_single_block(weights.single_blocks[0], torch.cat((T["img_after_double"], T["txt_after_double"])).cuda(), T["vec"].cuda(),
                                 Lt, cu, Li + Lt, freqs)
    f3, m3 = _bad(x, T["x_after_single"])
    print(f"hunyuan single block: bad {f3:.2e} max {m3:.4f}")
    assert f3 < 2e-3 and m3 < 0.13


def test_two_segment_varlen_vs_oracle_on_gpu():
    """1 double + 2 single blocks, 1200 image tokens + 256 text tokens of which 77 are valid: cu_seqlens = [0, 1277, 1456]."""
    hidden, mlp, heads = 3072, 12288, 24
    W = HO.synth_weights(1, 2, hidden, mlp, seed=3, device="cuda")
    img, txt, vec, cu, freqs = HO.synth_inputs(1200, 256, 77, hidden, seed=4, device="cuda")
    ref = HO.infer_blocks(W, 1, 2, img.clone(), txt.clone(), vec, cu, freqs, heads, attn="flash_attn2")
    weights, infer = _build(W, 1, 2)
    out, _ = infer.infer(weights, img.clone(), txt.clone(), vec, torch.tensor(cu, dtype=torch.int32), 1456, freqs)
    torch.cuda.synchronize()
    f, m = _bad(out, ref)
    from oracle.wan_oracle import psnr
    p = psnr(out, ref)
    print(f"hunyuan 1+2 blocks, two varlen segments: bad {f:.2e} max {m:.4f} psnr {p:.1f} dB")
    # three chained blocks with |x| up to ~9 (bf16 ulp 0.03-0.06 there): a two-ulp drift exceeds rtol = atol = 1e-2 on ~1 % of the
    # elements between ANY two bf16 implementations (flash-attn vs this FMHA, cuBLAS vs this GEMM); single blocks are pinned at
    # 2e-3 against the real reference above.  Gate: <= 2 % outside the tolerance, no element beyond 0.13, PSNR >= 45 dB.
    assert f < 2e-2 and m < 0.13 and p > 45


def test_hunyuan_pre_and_post_infer_vs_reference_fixture(golden_dir, record):
    """host/hunyuan_prepost.py on the CUDA kernels vs the fixture of the REAL HunyuanPreInfer methods (time / guidance / vector embedders,
    patch embedding) and the REAL HunyuanPostInfer (adaLN + FP32 projection + unpatchify)."""
    import os

    from safetensors import safe_open

    from lightx2v_b200.host.hunyuan_prepost import HunyuanPostInfer, HunyuanPreInfer
    from oracle import hunyuan_oracle as HO

    with safe_open(os.path.join(golden_dir, "hunyuan_prepost.safetensors"), framework="pt") as f:
        T = {k: f.get_tensor(k).cuda() for k in f.keys()}
        meta = f.metadata()
    W = {k: v.cuda() for k, v in HO.synth_prepost_weights(int(meta["hidden"]), seed=int(meta["weights_seed"])).items()}
    cfg = dict(task="t2v", mm_config={})
    pre = HunyuanPreInfer(cfg)

    def close(got, ref, name, rtol=1e-2, atol=1e-2):
        err = (got.float() - ref.float()).abs()
        bad = (err > atol + rtol * ref.float().abs()).float().mean().item()
        record(**{name + "_bad_frac": bad, name + "_max": err.max().item()})
        assert bad < 1e-3, (name, bad, err.max().item())

    close(pre.infer_time_in(W, T["t"][0]), T["time_out"], "time_in")
    close(pre.infer_guidance_in(W, T["guidance"]), T["guidance_out"], "guidance_in")
    close(pre.infer_vector_in(W, T["text_states_2"]), T["vector_out"], "vector_in")
    close(pre.infer_img_in(W, T["latents"]), T["img_out"], "img_in")

    class Sched:
        latents = T["latents"]

    post = HunyuanPostInfer(cfg)
    post.set_scheduler(Sched())
    out = post.infer(W, T["img"], T["vec"])
    assert out.dtype == torch.float32 and out.shape == T["post_out"].shape
    close(out, T["post_out"], "post_infer")


def test_i2v_token_replace_blocks_vs_reference_fixture(golden_dir, record):
    """HunyuanVideo i2v: the first rows of the image stream (the conditioning frame's tokens) take their modulation from
    mod(silu(token_replace_vec)); CUDA path vs the fixture of the REAL HunyuanTransformerInfer (one double + one single block)."""
    with safe_open(os.path.join(golden_dir, "hunyuan_blocks_i2v_small.safetensors"), framework="pt") as f:
        T = {k: f.get_tensor(k) for k in f.keys()}
        meta = f.metadata()
    hidden, mlp, first = int(meta["hidden"]), int(meta["mlp"]), int(meta["first_frame_tokens"])
    W = HO.synth_weights(1, 1, hidden, mlp, seed=int(meta["weights_seed"]))
    weights, infer = _build(W, 1, 1)
    Li, Lt = T["img"].shape[0], T["txt"].shape[0]
    cu = [0, Li + int(meta["txt_valid"]), Li + Lt]
    freqs = (T["cos"].cuda(), T["sin"].cuda())
    trv = T["token_replace_vec"].cuda()
    img, txt = infer.infer_double_block(weights.double_blocks[0], T["img"].cuda().clone(), T["txt"].cuda().clone(), T["vec"].cuda(), cu, Li + Lt, freqs, trv, first)
    f1, m1 = _bad(img, T["img_after_double"])
    f2, m2 = _bad(txt, T["txt_after_double"])
    x = infer.infer_single_block(weights.single_blocks[0], torch.cat((T["img_after_double"], T["txt_after_double"])).cuda(), T["vec"].cuda(), Lt, cu, Li + Lt,
                                 freqs, trv, first)
    f3, m3 = _bad(x, T["x_after_single"])
    record(double_img_bad_frac=f1, double_img_max=m1, double_txt_bad_frac=f2, single_bad_frac=f3, single_max=m3)
    assert f1 < 2e-3 and f2 < 2e-3 and f3 < 2e-3 and max(m1, m2, m3) < 0.13
    # and the t2v path of the same weights differs (the replacement is not a no-op; the later rows change too, through attention)
    img_t2v, _ = infer.infer_double_block(weights.double_blocks[0], T["img"].cuda().clone(), T["txt"].cuda().clone(), T["vec"].cuda(), cu, Li + Lt, freqs)
    assert _bad(img_t2v[:first], T["img_after_double"][:first])[0] > 0.1
