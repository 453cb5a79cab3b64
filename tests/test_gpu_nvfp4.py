"""GPU parity of the w4a4 NVFP4 path: the quantiser against the oracle (= the reference's fake_quant.py golden model) BIT FOR BIT
(packed e2m1 bytes up to the sign of zero, ue4m3 scale bytes in the 128x4 MMA layout), and the block-scaled tcgen05 GEMM against
the reference's own acceptance test: dequantise both operands to fp32, matmul, add bias, assert_close(atol=1e-1, rtol=1e-1)
(lightx2v_kernel/test/nvfp4_nvfp4/test_bench1.py:57-138) - tightened here to the bf16 output rounding, since with identical
quantised operands the only differences are fp32 summation order and the final bf16 cast."""
import os

import pytest
import torch
from safetensors import safe_open

from oracle import nvfp4_oracle as NV

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from lightx2v_b200 import lib as L

    L.load()
    return L


def _norm_zero_sign(packed):
    """-0 (nibble 0x8) and +0 (0x0) are the same e2m1 value; the golden model produces floats, so normalise before comparing bytes."""
    lo, hi = packed & 0x0F, packed >> 4
    lo = torch.where(lo == 8, torch.zeros_like(lo), lo)
    hi = torch.where(hi == 8, torch.zeros_like(hi), hi)
    return lo | (hi << 4)


@pytest.mark.parametrize("rows,K", [(200, 256), (128, 64), (1, 512), (777, 1024)])
def test_quant_nvfp4_bit_exact_vs_oracle(lib, rows, K):
    g = torch.Generator().manual_seed(rows + K)
    x = (torch.randn(rows, K, generator=g) * 2.3).to(torch.bfloat16)
    if rows > 8:
        x[5, 32:48] = 0
        x[7, 0] = 55.0
    gs = NV.global_scale_for(x)
    ref_q, ref_sf = NV.scaled_fp4_quant(x, gs)
    q, sf = lib.quant_nvfp4(x.cuda(), gs.reshape(1).cuda())
    assert q.shape == ref_q.shape and sf.shape == ref_sf.shape
    assert torch.equal(sf.cpu(), ref_sf)
    assert torch.equal(_norm_zero_sign(q.cpu()), _norm_zero_sign(ref_q))
    gs_dev, _ = lib.nvfp4_act_scale(x.cuda())
    assert torch.equal(gs_dev.cpu(), gs.reshape(1))


def test_quant_matches_reference_fixture(lib, golden_dir):
    with safe_open(os.path.join(golden_dir, "nvfp4_quant_small.safetensors"), framework="pt") as f:
        a, gs_a, qa, sa = f.get_tensor("a"), f.get_tensor("gs_a"), f.get_tensor("qa"), f.get_tensor("sa")
    q, sf = lib.quant_nvfp4(a.cuda(), gs_a.cuda())
    assert torch.equal(NV.unpack_e2m1(q.cpu()), qa)                              # values on the e2m1 grid (-0 == 0)
    assert torch.equal(NV.unswizzle_sf(sf.cpu(), *a.shape).view(torch.float8_e4m3fn).float(), sa)


@pytest.mark.parametrize("M,N,K,block_n", [(200, 96, 256, 128), (128, 128, 256, 128), (300, 384, 1024, 128), (300, 384, 1024, 256),
                                           (1000, 5120, 5120, 128), (1000, 5120, 5120, 256), (130, 1000, 768, 128), (200, 136, 320, 128)])
def test_gemm_nvfp4_vs_dequantised_matmul(lib, M, N, K, block_n):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = (torch.randn(M, K, generator=g, device="cuda") * 1.3).to(torch.bfloat16)
    b = (torch.randn(N, K, generator=g, device="cuda") * 0.04).to(torch.bfloat16)
    bias = torch.randn(N, generator=g, device="cuda").to(torch.bfloat16)
    gs_b = NV.global_scale_for(b.cpu()).reshape(1).cuda()
    gs_a, alpha = lib.nvfp4_act_scale(a, gs_b)
    aq, sfa = lib.quant_nvfp4(a, gs_a)
    bq, sfb = lib.quant_nvfp4(b, gs_b)
    out = lib.gemm_nvfp4(aq, bq, sfa, sfb, alpha, bias, block_n=block_n)
    ref = NV.scaled_fp4_mm(aq.cpu(), bq.cpu(), sfa.cpu(), sfb.cpu(), gs_a.cpu()[0], gs_b.cpu()[0], bias.cpu())
    torch.testing.assert_close(out.float().cpu(), ref, atol=1e-1, rtol=1e-1)     # the reference's own bar (test_bench1.py:138)
    err = (out.float().cpu() - ref).abs()
    assert err.max() <= 2e-2 + 8e-3 * ref.abs().max(), (float(err.max()), float(ref.abs().max()))
    # and the quantisation itself is sane: PSNR of the w4a4 product against the bf16 product
    full = a.float() @ b.float().t() + bias.float()
    mse = (out.float() - full).pow(2).mean()
    assert 10 * torch.log10(full.abs().max() ** 2 / mse) > 25


def test_gemm_nvfp4_fixture_and_epilogues(lib, golden_dir):
    with safe_open(os.path.join(golden_dir, "nvfp4_quant_small.safetensors"), framework="pt") as f:
        T = {k: f.get_tensor(k) for k in f.keys()}
    a, b, bias = T["a"].cuda(), T["b"].cuda(), T["bias"].cuda()
    gs_a, gs_b = T["gs_a"].cuda(), T["gs_b"].cuda()
    aq, sfa = lib.quant_nvfp4(a, gs_a)
    bq, sfb = lib.quant_nvfp4(b, gs_b)
    alpha = (1.0 / (gs_a * gs_b)).float()
    out = lib.gemm_nvfp4(aq, bq, sfa, sfb, alpha, bias)
    assert (out.float().cpu() - T["out"]).abs().max() <= 2e-2 + 8e-3 * T["out"].abs().max()
    # residual epilogue: x + bf16(acc + bias)
    x0 = torch.randn(200, 96, device="cuda").to(torch.bfloat16)
    x = x0.clone()
    lib.gemm_nvfp4(aq, bq, sfa, sfb, alpha, bias, out=x, epilogue=lib.EPI_RESIDUAL)
    want = (x0.float() + out.float()).to(torch.bfloat16)
    assert (x.float() - want.float()).abs().max() <= 2 ** -6 * want.float().abs().max()


def test_nvfp4_block_psnr_vs_oracle_and_bf16_reference():
    """One 14B-width Wan block in w4a4 NVFP4 through the fused path: (a) against the oracle's fake-quant restatement (same quantised
    operands, fp32 matmul), (b) PSNR against the bf16 reference path (north_star: nvfp4 reports PSNR vs the bf16 reference)."""
    from oracle import wan_oracle as O
    from lightx2v_b200.host.ops import NVFP4_MM_KEY, quantize_checkpoint_nvfp4
    from lightx2v_b200.host.wan_infer import WanTransformerInfer
    from lightx2v_b200.host.wan_weights import WanTransformerWeights

    dim, heads, ffn, grid = 5120, 40, 13824, (10, 6, 10)
    S = grid[0] * grid[1] * grid[2]
    W = O.synth_block_weights(1, dim, ffn, seed=1, device="cuda")
    x, embed0, context = O.synth_block_inputs(S, dim, seed=2, device="cuda")
    freqs = O.wan_freqs_table(128)
    ref_bf16 = O.infer_blocks(W, 1, x.clone(), embed0, grid, freqs.cuda(), context, heads, attn="flash_attn2")
    ref_f4 = O.infer_blocks(O.quantize_checkpoint_nvfp4(W), 1, x.clone(), embed0, grid, freqs.cuda(), context, heads, attn="flash_attn2")
    cfg = dict(task="t2v", num_layers=1, num_heads=heads, dim=dim, ffn_dim=ffn, mm_config={"mm_type": NVFP4_MM_KEY})
    weights = WanTransformerWeights(cfg)
    weights.load(W)                                           # bf16 checkpoint on the GPU: quantised at load
    infer = WanTransformerInfer(cfg)
    out = infer.infer(weights, torch.tensor([grid]), None, x.clone(), embed0, None, freqs, context)
    torch.cuda.synchronize()
    p_impl, p_q, p_ref = O.psnr(out, ref_f4), O.psnr(out, ref_bf16), O.psnr(ref_f4, ref_bf16)
    print(f"nvfp4 block: PSNR vs fake-quant oracle {p_impl:.1f} dB; vs bf16 reference {p_q:.1f} dB (oracle nvfp4 vs bf16: {p_ref:.1f} dB)")
    assert p_impl > 35.0          # same quantised arithmetic up to bf16 rounding of intermediates feeding the next dynamic quantisation
    assert p_q > p_ref - 1.0      # no worse than the golden model's own w4a4 error against bf16
    # the offline converter emits exactly what load-time quantisation produces, and a packed checkpoint loads unchanged
    names = [k for k in W if k.endswith(".weight") and W[k].dim() == 2 and ("attn." in k or "ffn." in k) and "norm" not in k]
    Wp = quantize_checkpoint_nvfp4(W, names)
    w2 = WanTransformerWeights(cfg)
    w2.load(Wp)
    w2.to_cuda()
    a, b = w2.blocks[0].compute_phases[3].ffn_0, weights.blocks[0].compute_phases[3].ffn_0
    assert torch.equal(a.weight, b.weight) and torch.equal(a.weight_scale, b.weight_scale) and torch.equal(a.weight_global_scale, b.weight_global_scale)
    sd = a.state_dict()
    assert sd[a.weight_name].dtype == torch.uint8 and a.weight_scale_name in sd and a.weight_global_scale_name in sd


def test_full_size_properties_nvfp4(lib):
    """BASELINE sizes (75 600 tokens x 5120 -> 5120), where the oracle cannot run in full: (a) quantiser rows sampled across the whole
    tensor are bit-identical to the oracle, incl. the scale-factor layout at large row indices; (b) the two tile shapes of the GEMM
    (BLOCK_N 128 / 256, same K order) agree bit for bit; (c) doubling alpha doubles the output exactly (power-of-two linearity);
    (d) sampled output rows match the dequantised fp32 matmul."""
    M, N, K = 75600, 5120, 5120
    g = torch.Generator(device="cuda").manual_seed(123)
    a = (torch.randn(M, K, generator=g, device="cuda") * 1.1).to(torch.bfloat16)
    b = (torch.randn(N, K, generator=g, device="cuda") * 0.03).to(torch.bfloat16)
    gs_b, _ = lib.nvfp4_act_scale(b)
    gs_a, alpha = lib.nvfp4_act_scale(a, gs_b)
    assert torch.equal(gs_a.cpu(), NV.global_scale_for(a.cpu()).reshape(1)), "global scale"
    aq, sfa = lib.quant_nvfp4(a, gs_a)
    bq, sfb = lib.quant_nvfp4(b, gs_b)
    rows = torch.tensor([0, 1, 127, 128, 4095, 37777, 75519, 75599])
    ref_q, ref_sf = NV.quant(a[rows.cuda()].cpu(), gs_a.cpu()[0])
    assert torch.equal(NV.unpack_e2m1(aq[rows.cuda()].cpu()), ref_q), "sampled e2m1 rows"
    sf_lin = NV.unswizzle_sf(sfa.cpu(), M, K)
    assert torch.equal(sf_lin[rows].view(torch.float8_e4m3fn).float(), ref_sf), "sampled scale rows"
    o128 = lib.gemm_nvfp4(aq, bq, sfa, sfb, alpha, block_n=128)
    o256 = lib.gemm_nvfp4(aq, bq, sfa, sfb, alpha, block_n=256)
    assert torch.equal(o128, o256), f"tile shapes differ: {(o128.float() - o256.float()).abs().max().item()} on {(o128 != o256).float().mean().item():.2e} of the elements"
    o2 = lib.gemm_nvfp4(aq, bq, sfa, sfb, alpha * 2, block_n=128)
    assert torch.equal(o2.float(), o128.float() * 2), "alpha linearity"
    ref = NV.scaled_fp4_mm(aq[rows.cuda()].cpu(), bq.cpu(), sf_rows_swizzled(sf_lin[rows], K), sfb.cpu(), gs_a.cpu()[0], gs_b.cpu()[0])
    err = (o128[rows.cuda()].float().cpu() - ref).abs()
    assert err.max() <= 2e-2 + 8e-3 * ref.abs().max(), (float(err.max()), float(ref.abs().max()))


def sf_rows_swizzled(sf_rows_linear, K):
    """Re-swizzle a few linear scale rows so the oracle's dequantiser (which expects the MMA layout) can consume them."""
    return NV.swizzle_sf(sf_rows_linear, sf_rows_linear.shape[0], K)
