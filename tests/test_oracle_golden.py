"""Pins oracle/wan_oracle.py against fixtures produced by the REAL reference classes (oracle/gen_golden.py)."""
import os

import pytest
import torch
from safetensors import safe_open

from oracle import wan_oracle as O


def _load(path):
    with safe_open(path, framework="pt") as f:
        return {k: f.get_tensor(k) for k in f.keys()}, f.metadata()


@pytest.mark.parametrize("name", ["wan13b_t2v_2blocks", "wan13b_i2v_1block"])
def test_oracle_matches_reference_fixture(golden_dir, name):
    torch.set_num_threads(8)
    T, meta = _load(os.path.join(golden_dir, name + ".safetensors"))
    dim, heads, ffn, L, task = int(meta["dim"]), int(meta["heads"]), int(meta["ffn"]), int(meta["layers"]), meta["task"]
    W = O.synth_block_weights(L, dim, ffn, task=task, seed=int(meta["weights_seed"]))
    grid = T["grid"].tolist()
    freqs = O.wan_freqs_table(dim // heads)

    # RoPE probe: bit exact
    fi = O.compute_freqs(dim // heads // 2, grid, freqs)
    assert torch.equal(O.apply_rotary_emb(T["probe.rope_in"], fi), T["probe.rope_out"])

    # per-function probes on block 0, bit exact (same torch CPU kernels, same op order as the reference)
    x = T["x_in"].clone()
    pre = "blocks.0."
    sh, sc, ga, csh, csc, cga = O.infer_modulation(W, pre, T["embed0"])
    y = O.infer_self_attn(W, pre, x, fi, sh, sc, heads)
    assert torch.equal(y, T["probe.self_attn_y"])
    x2, attn_out = O.infer_cross_attn(W, pre, x, T["context"], y, ga, heads, task)
    assert torch.equal(x2, T["probe.x_after_cross"])
    assert torch.equal(attn_out, T["probe.cross_attn_out"])
    y_ffn = O.infer_ffn(W, pre, x2, attn_out, csh, csc)
    assert torch.equal(y_ffn, T["probe.ffn_y"])

    # whole stack
    out = O.infer_blocks(W, L, T["x_in"].clone(), T["embed0"], grid, freqs, T["context"], heads, task)
    assert torch.equal(out, T["x_out"])


def test_freqs_dist_padding_is_identity():
    freqs = O.wan_freqs_table(128)
    fi = O.compute_freqs_dist(10, 64, (1, 3, 5), freqs, world_size=2, rank=1)   # 15 tokens padded to 20
    assert fi.shape == (10, 1, 64)
    assert torch.all(fi[5:] == 1)
    full = O.compute_freqs(64, (1, 3, 5), freqs)
    assert torch.equal(fi[:5], full[10:15])


def test_block_flops_matches_survey():
    # 163.1 TFLOP per block at 14B / 720p
    fl = O.block_flops(75600, 5120, 13824)
    assert abs(fl / 1e12 - 163.1) < 0.2


def test_vae_oracle_matches_reference_fixture(golden_dir):
    """oracle/vae_oracle.py vs the REAL WanVAE_.decode (per-frame causal-cache loop), bit for bit."""
    from oracle import vae_oracle as V

    T, _ = _load(os.path.join(golden_dir, "wan_vae_decode_small.safetensors"))
    out = V.vae_decode(V.synth_vae_weights(0), T["zs"])
    assert out.shape == (1, 3, 9, 64, 64)
    # fp32 convolutions: oneDNN's summation order depends on thread count / primitive cache, so exact equality holds only
    # run-to-run in the same process state; tolerance 2e-5 absolute on outputs in [-1, 1] (observed: 0 .. 4e-6)
    err = (out - T["images"]).abs().max().item()
    assert err <= 2e-5, err
    assert V.count_causal_convs(V.decoder_layout()[1]) == 33          # 33 CausalConv3d in the decoder


def test_hunyuan_oracle_matches_reference_fixture(golden_dir):
    """oracle/hunyuan_oracle.py vs the REAL HunyuanTransformerInfer double + single block at width 3072, bit for bit."""
    from oracle import hunyuan_oracle as HO

    torch.set_num_threads(8)
    T, meta = _load(os.path.join(golden_dir, "hunyuan_blocks_small.safetensors"))
    hidden, mlp, heads = int(meta["hidden"]), int(meta["mlp"]), int(meta["heads"])
    W = HO.synth_weights(1, 1, hidden, mlp, seed=int(meta["weights_seed"]))
    L_img, L_txt = T["img"].shape[0], T["txt"].shape[0]
    cu = [0, L_img + int(meta["txt_valid"]), L_img + L_txt]
    freqs = (T["cos"], T["sin"])
    img1, txt1 = HO.infer_double_block(W, 0, T["img"].clone(), T["txt"].clone(), T["vec"], cu, freqs, heads)
    assert torch.equal(img1, T["img_after_double"]) and torch.equal(txt1, T["txt_after_double"])
    x2 = HO.infer_single_block(W, 0, torch.cat((img1, txt1)), T["vec"], L_txt, cu, freqs, heads, hidden)
    assert torch.equal(x2, T["x_after_single"])


def test_hunyuan_vae_oracle_matches_reference_fixture(golden_dir):
    """oracle/hunyuan_vae_oracle.py vs the REAL AutoencoderKLCausal3D tiled decode (2 temporal x 2 x 2 spatial tiles, all three blends)
    and one raw un-tiled decoder tile; fp32 convolutions -> tolerance 2e-5 (observed 1.5e-6 / 3.2e-6)."""
    from oracle import hunyuan_vae_oracle as HV

    T, meta = _load(os.path.join(golden_dir, "hunyuan_vae_decode_small.safetensors"))
    cfg = dict(HV.HUNYUAN_VAE_CFG, block_out_channels=tuple(int(c) for c in meta["block_out_channels"].split(",")),
               sample_size=int(meta["sample_size"]), sample_tsize=int(meta["sample_tsize"]))
    W = HV.synth_vae_weights(cfg, seed=int(meta["weights_seed"]))
    with torch.no_grad():
        out = HV.decode(W, T["latents"], cfg)
        one = HV.tile_decode(W, (T["latents"] / cfg["scaling_factor"])[:, :, :3, :8, :8], cfg)
    assert out.shape == T["images"].shape == (1, 3, 21, 96, 80)
    assert (out - T["images"]).abs().max().item() <= 2e-5
    assert (one - T["tile_raw"]).abs().max().item() <= 5e-5
    # frame-causal mask of the mid-block attention: token of frame i sees frames <= i only
    m = HV.causal_frame_mask(3, 2, torch.float32, "cpu")
    assert m.shape == (6, 6) and m[0, 2] == float("-inf") and m[2, 1] == 0 and m[5, 0] == 0 and m[3, 4] == float("-inf")


def test_nvfp4_oracle_matches_reference_fixture(golden_dir):
    """oracle/nvfp4_oracle.py vs the reference's own fake_quant.py golden model, bit for bit; layout helpers round-trip."""
    from oracle import nvfp4_oracle as NV

    T, _ = _load(os.path.join(golden_dir, "nvfp4_quant_small.safetensors"))
    gs_a, gs_b = T["gs_a"][0], T["gs_b"][0]
    qa, sa = NV.quant(T["a"], gs_a)
    qb, sb = NV.quant(T["b"], gs_b)
    assert torch.equal(qa, T["qa"]) and torch.equal(sa, T["sa"]) and torch.equal(qb, T["qb"]) and torch.equal(sb, T["sb"])
    assert int((sa == 0).sum()) == 1                                            # the all-zero group
    pa, sfa = NV.scaled_fp4_quant(T["a"], gs_a)
    pb, sfb = NV.scaled_fp4_quant(T["b"], gs_b)
    assert pa.shape == (200, 128) and sfa.shape == (256, 16) and sfb.shape == (128, 16)
    assert torch.equal(NV.unpack_e2m1(pa), qa) and torch.equal(NV.unswizzle_sf(sfa, 200, 256).view(torch.float8_e4m3fn).float(), sa)
    # scale-factor layout: byte offset formula of nvfp4_quant_kernels_sm120.cu:127-152
    m, g = 137, 9
    off = ((m // 128) * (256 // 64) + g // 4) * 512 + (m % 32) * 16 + ((m % 128) // 32) * 4 + g % 4
    assert sfa.reshape(-1)[off] == sa.to(torch.float8_e4m3fn).view(torch.uint8)[m, g]
    out = NV.scaled_fp4_mm(pa, pb, sfa, sfb, gs_a, gs_b, T["bias"])
    assert torch.allclose(out, T["out"], rtol=1e-5, atol=1e-4)


def test_causvid_oracle_matches_reference_fixture(golden_dir):
    """oracle.wan_oracle.infer_blocks_causvid vs the REAL WanTransformerInferCausVid: three chunks through two blocks with the KV cache
    growing, bit for bit (outputs of every chunk and the final K cache of block 0)."""
    torch.set_num_threads(8)
    T, meta = _load(os.path.join(golden_dir, "wan13b_causvid_2blocks.safetensors"))
    dim, heads, ffn, L = int(meta["dim"]), int(meta["heads"]), int(meta["ffn"]), int(meta["layers"])
    chunks, ft = int(meta["chunks"]), int(meta["frame_tokens"])
    grid = tuple(int(v) for v in meta["grid"].split(","))
    W = O.synth_block_weights(L, dim, ffn, seed=int(meta["weights_seed"]))
    freqs = O.wan_freqs_table(dim // heads)
    caches = [{"k": torch.zeros(chunks * ft, heads, dim // heads, dtype=torch.bfloat16), "v": torch.zeros(chunks * ft, heads, dim // heads, dtype=torch.bfloat16)}
              for _ in range(L)]
    for c in range(chunks):
        out = O.infer_blocks_causvid(W, L, T[f"x_in.{c}"].clone(), T[f"embed0.{c}"], grid, freqs, T["context"], heads, caches, c * ft, (c + 1) * ft)
        assert torch.equal(out, T[f"x_out.{c}"]), c
    assert torch.equal(caches[0]["k"].reshape(chunks * ft, dim), T["k_cache.0"])


@pytest.mark.parametrize("task", ["t2v", "i2v"])
def test_prepost_oracle_matches_reference_fixture(golden_dir, task):
    """oracle.wan_oracle.pre_infer / post_infer vs the REAL WanPreInfer / WanPostInfer (A13), bit for bit."""
    torch.set_num_threads(8)
    T, meta = _load(os.path.join(golden_dir, "wan13b_prepost.safetensors"))
    dim = int(meta["dim"])
    g = lambda k: T[f"{task}.{k}"]       # noqa: E731
    W = O.synth_prepost_weights(dim, 36 if task == "i2v" else 16, task, seed=int(meta["weights_seed"]))
    t = g("timesteps")[int(meta["step_index"])].reshape(1)
    extra = dict(clip_fea=g("clip_encoder_out"), vae_encode_out=g("vae_encode_out")) if task == "i2v" else {}
    embed, grid, x, embed0, ctx = O.pre_infer(W, g("latents"), t, g("context"), dim, **extra)
    assert grid == tuple(int(v) for v in g("grid"))
    assert torch.equal(embed, g("embed")) and torch.equal(x, g("x")) and torch.equal(embed0, g("embed0")) and torch.equal(ctx, g("context_out"))
    noise = O.post_infer(W, g("x_blocks").clone(), embed, grid)
    assert torch.equal(noise, g("noise_pred"))


def test_cogvideox_oracle_matches_reference_fixture(golden_dir):
    """oracle/cogvideox_oracle.py vs the fixture of the REAL CogvideoxTransformerInfer + CogVideoXBlock classes: bit for bit."""
    from oracle import cogvideox_oracle as C

    torch.set_num_threads(8)
    T, meta = _load(os.path.join(golden_dir, "cogvideox_2blocks.safetensors"))
    layers, heads, hd, ff = int(meta["layers"]), int(meta["heads"]), int(meta["head_dim"]), int(meta["ff"])
    W = C.synth_weights(layers, heads * hd, ff, hd, seed=int(meta["weights_seed"]))
    rotary = (T["cos"], T["sin"])
    grid = [int(v) for v in meta["grid"].split(",")]
    r2 = C.rotary_table(*grid, head_dim=hd, seed=int(meta["rotary_seed"]))
    assert torch.equal(r2[0], T["cos"]) and torch.equal(r2[1], T["sin"])
    nh, ne, gate, _ = C.norm_mod(W, "transformer_blocks.0.", "norm1", T["hidden_in"].clone(), T["enc_in"].clone(), T["temb"])
    assert torch.equal(nh, T["probe.norm1_hidden"]) and torch.equal(ne, T["probe.norm1_enc"]) and torch.equal(gate, T["probe.gate"])
    ah, ae = C.attention(W, "transformer_blocks.0.", nh.clone(), ne.clone(), rotary, heads)
    assert torch.equal(ah, T["probe.attn_hidden"]) and torch.equal(ae, T["probe.attn_enc"])
    h, e = C.infer_blocks(W, layers, T["hidden_in"].clone(), T["enc_in"].clone(), T["temb"], rotary, heads)
    assert torch.equal(h, T["hidden_out"]) and torch.equal(e, T["enc_out"])


def test_hunyuan_prepost_oracle_matches_reference_fixture(golden_dir):
    """oracle/hunyuan_oracle.py pre/post-infer functions vs the REAL HunyuanPreInfer / HunyuanPostInfer methods: bit for bit."""
    from oracle import hunyuan_oracle as HO

    T, meta = _load(os.path.join(golden_dir, "hunyuan_prepost.safetensors"))
    W = HO.synth_prepost_weights(int(meta["hidden"]), seed=int(meta["weights_seed"]))
    assert torch.equal(HO.infer_time_in(W, T["t"][0]), T["time_out"])
    assert torch.equal(HO.infer_guidance_in(W, T["guidance"]), T["guidance_out"])
    assert torch.equal(HO.infer_vector_in(W, T["text_states_2"]), T["vector_out"])
    assert torch.equal(HO.infer_img_in(W, T["latents"]), T["img_out"])
    assert torch.equal(HO.post_infer(W, T["img"], T["vec"], T["latents"].shape), T["post_out"])
    mask = torch.zeros(1, 256, dtype=torch.int64)
    mask[0, :77] = 1
    assert HO.cu_seqlens(mask, 1000) == [0, 1077, 1256]


def test_vae_encoder_oracle_matches_reference_fixture(golden_dir):
    """oracle/vae_oracle.py:vae_encode vs the REAL WanVAE_.encode (fp32 CPU): bit for bit."""
    from oracle import vae_oracle as V

    torch.set_num_threads(8)
    T, _ = _load(os.path.join(golden_dir, "wan_vae_encode_small.safetensors"))
    out = V.vae_encode(V.synth_vae_encoder_weights(0), T["video"])
    assert torch.equal(out, T["mu"])


def test_hunyuan_i2v_token_replace_oracle_matches_reference_fixture(golden_dir):
    """oracle/hunyuan_oracle.py with token_replace_vec / first_frame_tokens vs the REAL HunyuanTransformerInfer i2v path, bit for bit."""
    from oracle import hunyuan_oracle as HO

    torch.set_num_threads(8)
    T, meta = _load(os.path.join(golden_dir, "hunyuan_blocks_i2v_small.safetensors"))
    hidden, mlp, heads, first = int(meta["hidden"]), int(meta["mlp"]), int(meta["heads"]), int(meta["first_frame_tokens"])
    W = HO.synth_weights(1, 1, hidden, mlp, seed=int(meta["weights_seed"]))
    L_img, L_txt = T["img"].shape[0], T["txt"].shape[0]
    cu = [0, L_img + int(meta["txt_valid"]), L_img + L_txt]
    freqs = (T["cos"], T["sin"])
    img1, txt1 = HO.infer_double_block(W, 0, T["img"].clone(), T["txt"].clone(), T["vec"], cu, freqs, heads, token_replace_vec=T["token_replace_vec"],
                                       first_frame_tokens=first)
    assert torch.equal(img1, T["img_after_double"]) and torch.equal(txt1, T["txt_after_double"])
    x2 = HO.infer_single_block(W, 0, torch.cat((img1, txt1)), T["vec"], L_txt, cu, freqs, heads, hidden, token_replace_vec=T["token_replace_vec"],
                               first_frame_tokens=first)
    assert torch.equal(x2, T["x_after_single"])
