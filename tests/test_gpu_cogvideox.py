"""CogVideoX block stack on the CUDA path (host/cogvideox_infer.py: head_dim-64 FMHA, per-head LayerNorm + pair RoPE kernel, fused
GEMM epilogues) vs the fixture produced by the REAL CogvideoxTransformerInfer classes (oracle/gen_golden.py:gen_cogvideox_fixture) and
vs the oracle restatement at the published width (48 heads x 64 = 3072, ff 12288).  Tolerance rtol = atol = 1e-2 (north_star); the
admitted one-ulp fraction is stated and the measured values are recorded."""
import os

import pytest
import torch
from safetensors import safe_open

from oracle import cogvideox_oracle as C

pytestmark = pytest.mark.gpu


def _load(path):
    with safe_open(path, framework="pt") as f:
        return {k: f.get_tensor(k) for k in f.keys()}, f.metadata()


def _bad_frac(got, ref, rtol=1e-2, atol=1e-2):
    got, ref = got.float().cpu(), ref.float().cpu()
    return ((got - ref).abs() > atol + rtol * ref.abs()).float().mean().item(), (got - ref).abs().max().item()


class _Sched:
    def __init__(self, rot):
        self.image_rotary_emb = rot


def _build(layers, heads, W):
    from lightx2v_b200.host.cogvideox_infer import CogvideoxTransformerInfer, CogvideoxTransformerWeights

    cfg = dict(num_layers=layers, transformer_num_layers=layers, transformer_num_attention_heads=heads, transformer_attention_head_dim=64)
    weights = CogvideoxTransformerWeights(cfg)
    weights.load_weights({k: v.cuda() for k, v in W.items()})
    return weights, CogvideoxTransformerInfer(cfg)


def test_ln_rope_heads64_kernel_vs_reference_ops():
    from lightx2v_b200 import lib

    g = torch.Generator(device="cuda").manual_seed(0)
    L, Lt, H = 300, 40, 6
    qkv = (torch.randn(L, 3, H, 64, generator=g, device="cuda") * 2).to(torch.bfloat16)
    wq, bq = (1 + 0.1 * torch.randn(64, generator=g, device="cuda")).to(torch.bfloat16), (0.1 * torch.randn(64, generator=g, device="cuda")).to(torch.bfloat16)
    wk, bk = (1 + 0.1 * torch.randn(64, generator=g, device="cuda")).to(torch.bfloat16), (0.1 * torch.randn(64, generator=g, device="cuda")).to(torch.bfloat16)
    cos, sin = C.rotary_table(2, 10, 13, 64, seed=1)
    cos, sin = cos.cuda(), sin.cuda()
    ref = []
    for i, (w, b) in enumerate(((wq, bq), (wk, bk))):
        x = qkv[:, i].transpose(0, 1).clone()                                    # [H, L, 64]
        x = torch.nn.functional.layer_norm(x, (64,), w, b, 1e-6)
        x[:, Lt:] = C.apply_rotary_emb(x[:, Lt:], (cos, sin))
        ref.append(x.transpose(0, 1))
    from lightx2v_b200.host.cogvideox_infer import rotary_pairs
    lib.ln_rope_heads64_(qkv[:, 0], wq, bq, qkv[:, 1], wk, bk, eps=1e-6, cos_sin=rotary_pairs((cos, sin)), rope_start=Lt)
    for i in range(2):
        f, m = _bad_frac(qkv[:, i], ref[i])
        assert f < 2e-3 and m < 0.04, (i, f, m)       # one bf16 ulp where torch's LN statistics round differently


def test_cogvideox_blocks_vs_reference_fixture(golden_dir, record):
    T, meta = _load(os.path.join(golden_dir, "cogvideox_2blocks.safetensors"))
    layers, heads, hd, ff = int(meta["layers"]), int(meta["heads"]), int(meta["head_dim"]), int(meta["ff"])
    W = C.synth_weights(layers, heads * hd, ff, hd, seed=int(meta["weights_seed"]))
    weights, infer = _build(layers, heads, W)
    infer.set_scheduler(_Sched((T["cos"].cuda(), T["sin"].cuda())))
    h, e = infer.infer(weights, T["hidden_in"].cuda(), T["enc_in"].cuda(), T["temb"].cuda())
    torch.cuda.synchronize()
    fh, mh = _bad_frac(h, T["hidden_out"])
    fe, me = _bad_frac(e, T["enc_out"])
    print(f"cogvideox 2 blocks: video bad_frac={fh:.3e} max={mh:.4f} | text bad_frac={fe:.3e} max={me:.4f}")
    record(video_bad_frac=fh, video_max=mh, text_bad_frac=fe, text_max=me)
    assert fh < 1e-3 and fe < 1e-3 and mh < 0.07 and me < 0.07
    # the reference's sub-steps on block 0
    blk = weights.blocks_weights[0]
    nh, ne, gate, _ = infer.cogvideox_norm1(blk, T["hidden_in"].cuda(), T["enc_in"].cuda(), T["temb"].cuda())
    assert _bad_frac(nh, T["probe.norm1_hidden"])[0] < 1e-3 and _bad_frac(ne, T["probe.norm1_enc"])[0] < 1e-3
    assert _bad_frac(gate, T["probe.gate"])[0] < 1e-3
    ah, ae = infer.cogvideox_attention(blk, T["probe.norm1_hidden"].cuda(), T["probe.norm1_enc"].cuda(), (T["cos"].cuda(), T["sin"].cuda()))
    assert _bad_frac(ah, T["probe.attn_hidden"])[0] < 1e-3 and _bad_frac(ae, T["probe.attn_enc"])[0] < 1e-3


def test_cogvideox_block_published_width_vs_oracle_on_gpu(record):
    """One block at the published width (48 heads x 64, ff 12288), 226 text + 4 x 12 x 20 = 960 video tokens, against the oracle restatement
    executed on the same GPU with torch ops (the reference's GPU path: torch.addmm, F.layer_norm, torch SDPA)."""
    heads, hd, ff, Lt, grid = 48, 64, 12288, 226, (4, 12, 20)
    dim, Lv = heads * hd, grid[0] * grid[1] * grid[2]
    W = C.synth_weights(1, dim, ff, hd, seed=5, device="cuda")
    hidden, enc, temb = C.synth_inputs(Lt, Lv, dim, seed=6, device="cuda")
    rot = tuple(t.cuda() for t in C.rotary_table(*grid, head_dim=hd, seed=2))
    want_h, want_e = C.infer_blocks(W, 1, hidden.clone(), enc.clone(), temb, rot, heads)
    weights, infer = _build(1, heads, W)
    infer.set_scheduler(_Sched(rot))
    h, e = infer.infer(weights, hidden, enc, temb)
    torch.cuda.synchronize()
    fh, mh = _bad_frac(h, want_h)
    fe, me = _bad_frac(e, want_e)
    record(video_bad_frac=fh, video_max=mh, text_bad_frac=fe, text_max=me)
    assert fh < 1e-3 and fe < 1e-3 and mh < 0.07 and me < 0.07, (fh, mh, fe, me)
