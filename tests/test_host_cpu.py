"""CPU-only tests: C-ABI surface (header <-> exported symbols <-> ctypes table), host-side mirror of the reference
operator interface, RoPE table construction, and the rule that the product path has no CPU fallback."""
import ctypes
import os
import re

import pytest
import torch

from oracle import wan_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as g

    if not os.path.exists(g.LIB_PATH):
        g.build()
    return ctypes.CDLL(g.LIB_PATH)


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "b200_dit.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_ctypes_binds():
    from lightx2v_b200 import lib

    assert _header_symbols() == sorted(lib.SIGNATURES.keys())


def test_library_exports_every_declared_symbol(built_lib):
    for sym in _header_symbols():
        assert hasattr(built_lib, sym), sym
    built_lib.b200_version.restype = ctypes.c_int
    assert built_lib.b200_version() >= 100


def test_argument_validation_without_gpu(built_lib):
    """Shape/alignment checks run before any CUDA call, so they are testable on the CPU box."""
    f = built_lib.b200_gemm_bf16
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_void_p, ctypes.c_int64] * 3 + [ctypes.c_void_p] * 2 + [ctypes.c_int64] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
    built_lib.b200_last_error.restype = ctypes.c_char_p
    assert f(None, 8, None, 8, None, 8, None, None, 8, 8, 8, 0, 0, 0, None) == -1
    assert b"null" in built_lib.b200_last_error()
    assert f(16, 100, 16, 100, 16, 8, None, None, 8, 8, 100, 0, 0, 0, None) == -1     # K % 8 != 0
    assert b"multiples of 8" in built_lib.b200_last_error()
    assert f(16, 64, 16, 64, 16, 64, None, None, 8, 64, 64, 2, 0, 0, None) == -1       # gate epilogue without gate
    g = built_lib.b200_fmha_fwd_d128
    g.restype = ctypes.c_int
    g.argtypes = [ctypes.c_void_p, ctypes.c_int64] * 4 + [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p]
    assert g(16, 128, 16, 128, 16, 128, 16, 128, 0, 5, 1, 1.0, None) == -1             # empty segment
    assert g(16, 100, 16, 128, 16, 128, 16, 128, 4, 4, 1, 1.0, None) == -1             # stride not multiple of 8


def test_no_cpu_fallback():
    from lightx2v_b200 import lib

    x = torch.zeros(8, 64, dtype=torch.bfloat16)
    with pytest.raises(lib.B200Error):
        lib.gemm_bf16(x, x)
    with pytest.raises(lib.B200Error):
        lib.ln_modulate(x)
    with pytest.raises(lib.B200Error):
        lib.fmha(torch.zeros(4, 1, 128, dtype=torch.bfloat16), torch.zeros(4, 1, 128, dtype=torch.bfloat16), torch.zeros(4, 1, 128, dtype=torch.bfloat16))


def test_product_code_never_imports_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "lightx2v_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+\"[^\"]*oracle", src, flags=re.M), f"{fn} uses the oracle"
                assert "oracle" not in src, f"{fn} mentions the oracle"


def test_registry_semantics_match_reference():
    from lightx2v_b200.host.registry import Register

    R = Register()

    @R("a")
    class A:  # noqa
        pass

    assert R["a"] is A and "a" in R
    with pytest.raises(Exception):
        R.register(A, key="a")                 # duplicate keys raise (registry_factory.py:19-20)
    R["a"] = int                               # __setitem__ overrides silently (registry_factory.py:25-26)
    assert R["a"] is int


def test_weight_tree_keys_and_state_dict_round_trip():
    from lightx2v_b200.host.wan_weights import WanTransformerWeights

    dim, ffn = 256, 512
    for task in ("t2v", "i2v"):
        W = O.synth_block_weights(2, dim, ffn, task=task, seed=5)
        cfg = dict(task=task, num_layers=2, num_heads=2, dim=dim, ffn_dim=ffn, mm_config={})
        tree = WanTransformerWeights(cfg)
        tree.load(W)
        blk = tree.blocks[1]
        assert [type(p).__name__ for p in blk.compute_phases] == ["WanModulation", "WanSelfAttention", "WanCrossAttention", "WanFFN"]
        sa = blk.compute_phases[1]
        assert sa.self_attn_q.weight.shape == (dim, dim) and not sa.self_attn_q.weight.is_contiguous()   # [K,N] view like mm_weight.py:76
        assert sa.self_attn_q.weight_nk.is_contiguous()
        sd = tree.state_dict()
        assert set(sd.keys()) == set(W.keys())
        for k in W:
            assert torch.equal(sd[k], W[k]), k
        assert tree.calculate_size() == sum(v.numel() * 2 for v in W.values())


def test_rope_cos_sin_matches_reference_tables():
    from lightx2v_b200.host.wan_infer import rope_cos_sin

    freqs = O.wan_freqs_table(128)
    grid = (3, 4, 5)
    t = rope_cos_sin(grid, freqs, 128)
    fi = O.compute_freqs(64, grid, freqs).reshape(60, 64)
    assert torch.equal(t[..., 0], fi.real.float()) and torch.equal(t[..., 1], fi.imag.float())
    # sharded (Ulysses) variant == compute_freqs_dist: pad rows are the identity rotation
    for rank in range(4):
        ts = rope_cos_sin(grid, freqs, 128, rows=16, row_offset=rank * 16)
        fd = O.compute_freqs_dist(16, 64, grid, freqs, 4, rank).reshape(16, 64)
        assert torch.equal(ts[..., 0], fd.real.float()) and torch.equal(ts[..., 1], fd.imag.float())


def test_fmha_op_segment_bounds():
    from lightx2v_b200.host.ops import FmhaWeightB200

    op = FmhaWeightB200()
    assert op._bounds(None, 7) == [0, 7]
    cu = torch.tensor([0, 5, 9], dtype=torch.int32)
    assert op._bounds(cu, 9) == [0, 5, 9]
    assert op._bounds(cu, 9) is op._bounds(cu, 9)              # cached per tensor object: no repeated device->host read
    cu[1] = 4                                                  # an in-place edit bumps _version and invalidates the entry
    assert op._bounds(cu, 9) == [0, 4, 9]
    assert op._bounds([0, 3], 3) == [0, 3]


def test_unipc_scheduler_matches_reference_fixture(golden_dir):
    """host/wan_scheduler.py vs the REAL WanScheduler run in the build container (oracle/gen_golden.py:gen_scheduler_fixture)."""
    from safetensors import safe_open

    from lightx2v_b200.host.wan_scheduler import WanScheduler

    with safe_open(os.path.join(golden_dir, "wan_scheduler_unipc.safetensors"), framework="pt") as f:
        T = {k: f.get_tensor(k) for k in f.keys()}
    cfg = dict(infer_steps=20, sample_shift=5.0, seed=42, target_shape=(16, 3, 8, 8), patch_size=(1, 2, 2))
    sch = WanScheduler(cfg, device="cpu")
    sch.prepare()
    assert torch.equal(sch.latents, T["latents_0"])
    assert torch.equal(sch.timesteps, T["timesteps"]) and torch.equal(sch.sigmas, T["sigmas"])
    for i in range(8):
        sch.step_pre(i)
        assert torch.equal(sch.latents, T[f"latents_pre_{i}"])
        sch.noise_pred = T[f"noise_pred_{i}"]
        sch.step_post()
        ref = T[f"latents_post_{i}"]
        assert sch.latents.dtype == ref.dtype
        assert torch.allclose(sch.latents, ref, rtol=1e-5, atol=1e-5), (i, (sch.latents - ref).abs().max())


def test_device_resident_scheduler_matches_the_eager_one_and_the_reference_fixture(golden_dir):
    """host/wan_graph.py:WanSchedulerDevice (coefficient table + static in-place state, the scheduler the per-step CUDA graphs capture) vs
    the eager WanScheduler and the REAL reference scheduler's fixture, step for step on the CPU: same latents bit for bit against the eager
    class (every step kind occurs: first, second, steady state, last), fixture tolerance as in the test above."""
    from safetensors import safe_open

    from lightx2v_b200.host.wan_graph import WanSchedulerDevice
    from lightx2v_b200.host.wan_scheduler import WanScheduler

    with safe_open(os.path.join(golden_dir, "wan_scheduler_unipc.safetensors"), framework="pt") as f:
        T = {k: f.get_tensor(k) for k in f.keys()}
    cfg = dict(infer_steps=20, sample_shift=5.0, seed=42, target_shape=(16, 3, 8, 8), patch_size=(1, 2, 2))
    eager, dev = WanScheduler(cfg, device="cpu"), WanSchedulerDevice(cfg, device="cpu")
    eager.prepare()
    dev.prepare()
    assert torch.equal(dev.s_lat, T["latents_0"]) and dev.table.shape == (20, 12) and len(set(dev.kinds)) >= 3
    g = torch.Generator().manual_seed(7)
    for i in range(20):
        eager.step_pre(i)
        dev.load_step(i)
        dev.step_pre()
        assert torch.equal(dev.latents, eager.latents), i
        # the fixture holds the reference's first 8 steps; later steps use fresh noise predictions
        pred = T[f"noise_pred_{i}"] if i < 8 else torch.randn(eager.latents.shape, generator=g)
        eager.noise_pred = pred
        dev.noise_pred = pred.clone()
        eager.step_post()
        dev.step_post()
        assert torch.equal(dev.latents, eager.latents), (i, dev.kind(i), (dev.latents - eager.latents).abs().max())
        if i < 8:
            assert torch.allclose(dev.latents, T[f"latents_post_{i}"], rtol=1e-5, atol=1e-5)
    assert dev.kind(0) == (0, 1) and dev.kind(19)[1] == 1          # first step: no corrector, order 1; last step: predictor order 1


def test_vae_dist_strip_bounds_match_reference_rules():
    """WanVAE.decode_dist (vae.py:883-922): 160 latent columns over 8 ranks -> 20-column chunks, 1-column halo, 160-pixel crops."""
    from lightx2v_b200.host.wan_vae import WanVAEDecoderB200 as D

    total, world = 160, 8
    cols = []
    for r in range(world):
        lat, crop = D.dist_slices(total, world, r)
        n = len(range(*lat.indices(total)))
        assert n == 22
        px = len(range(*crop.indices(n * 8)))
        assert px == 160
        # the cropped pixels of rank r are exactly the pixels of latent columns [20 r, 20 (r + 1))
        first_lat = lat.indices(total)[0]
        first_px = crop.indices(n * 8)[0]
        assert first_lat * 8 + first_px == 20 * r * 8
        cols.append(px)
    assert sum(cols) == 1280


def test_teacache_decisions_match_reference_fixture(golden_dir):
    """host/wan_teacache.py decision logic vs the decisions recorded from the REAL WanTransformerInferTeaCaching.calculate_should_calc
    (tests/golden/wan_teacache_decisions.safetensors), both `use_ret_steps` modes, cond / uncond interleaved."""
    import json

    from safetensors import safe_open

    from lightx2v_b200.host.wan_teacache import WanTransformerInferTeaCaching

    with safe_open(os.path.join(golden_dir, "wan_teacache_decisions.safetensors"), framework="pt") as f:
        T = {k: f.get_tensor(k) for k in f.keys()}
        meta = f.metadata()
    steps = T["embeds"].shape[0]
    for mode in (True, False):
        cfg = dict(task="t2v", num_layers=1, num_heads=12, dim=1536, infer_steps=steps, enable_cfg=True, teacache_thresh=float(meta["thresh"]),
                   coefficients=json.loads(meta["coefficients"]), use_ret_steps=mode)
        ti = WanTransformerInferTeaCaching(cfg)
        got = []
        for i in range(steps):
            for cond in (True, False):
                ti.infer_conditional = cond
                got.append(ti.calculate_should_calc(T["embeds"][i], T["embed0s"][i]))
                ti.cnt += 1
        assert got == [bool(v) for v in T[f"decisions_ret{int(mode)}"]], mode


def test_block_driver_structs_match_header():
    """The ctypes mirrors of b200_wan_block_weights / b200_wan_block_args must list the header's fields in the header's order
    (a silent mismatch would hand the native driver the wrong pointers)."""
    import re

    from lightx2v_b200 import lib

    hdr = open(os.path.join(ROOT, "include", "b200_dit.h")).read()

    def fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.sub(r"[\s\*]", " ", part).split()[-1])
        return names

    assert fields("b200_wan_block_weights") == [n for n, _ in lib.WanBlockWeightsC._fields_]
    assert fields("b200_wan_block_args") == [n for n, _ in lib.WanBlockArgsC._fields_]
    import ctypes as C
    types = dict(lib.WanBlockArgsC._fields_)
    assert types["S"] is C.c_int64 and types["ctx_len"] is C.c_int64 and types["D"] is C.c_int and types["eps"] is C.c_float


def test_argument_validation_of_the_newer_entry_points(built_lib):
    """Same idea as test_argument_validation_without_gpu for the nvfp4, VAE and block-driver entry points: bad shapes / pointers are
    rejected with a message before any CUDA call."""
    C = ctypes
    built_lib.b200_last_error.restype = C.c_char_p
    q = built_lib.b200_quant_nvfp4
    q.restype = C.c_int
    q.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    assert q(16, 96, 4, 96, 16, 16, 48, 16, None) == -1                      # K % 64 != 0
    assert b"multiple of 64" in built_lib.b200_last_error()
    assert q(None, 64, 4, 64, 16, 16, 32, 16, None) == -1
    g = built_lib.b200_gemm_nvfp4
    g.restype = C.c_int
    g.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64] + [C.c_void_p] * 4 + [C.c_int64, C.c_void_p, C.c_void_p] + [C.c_int64] * 3 + [C.c_int] * 3 + [C.c_void_p]
    assert g(16, 32, 16, 32, 16, 16, None, 16, 64, None, None, 8, 64, 64, 0, 0, 0, None) == -1       # alpha missing
    assert g(16, 32, 16, 32, 16, 16, 16, 16, 64, None, None, 8, 64, 100, 0, 0, 0, None) == -1        # K % 64 != 0
    cv = built_lib.b200_conv3d_cl_padded
    cv.restype = C.c_int
    cv.argtypes = ([C.c_void_p] + [C.c_int64] * 3 + [C.c_int] * 3 + [C.c_void_p] * 3 + [C.c_int64] * 3 + [C.c_void_p] + [C.c_int64] * 3 + [C.c_int] * 6
                   + [C.c_void_p, C.c_int, C.c_void_p])
    taps = (C.c_int32 * 3)(0, 0, 0)
    assert cv(16, 8, 8, 8, 4, 4, 4, 16, None, 16, 8, 8, 8, None, 0, 0, 0, 2, 2, 2, 48, 64, 1, C.cast(taps, C.c_void_p), 0, None) == -1   # cin % 32 != 0
    assert b"multiple of 32" in built_lib.b200_last_error()
    gs = built_lib.b200_gn_stats_cl
    gs.restype = C.c_int
    gs.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    assert gs(16, 0, 128, 16, None) == -1                                     # empty tensor
    wb = built_lib.b200_wan_block_fwd
    wb.restype = C.c_int
    wb.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    from lightx2v_b200 import lib
    w, a = lib.WanBlockWeightsC(), lib.WanBlockArgsC()
    a.S, a.D, a.H, a.F, a.eps = 64, 1536, 11, 8960, 1e-6                      # D != 128 * H
    assert wb(C.byref(w), C.byref(a), 16, 1 << 30, None) == -1
    assert b"128" in built_lib.b200_last_error()
    built_lib.b200_wan_block_workspace_bytes.restype = C.c_int64
    built_lib.b200_wan_block_workspace_bytes.argtypes = [C.c_int64, C.c_int, C.c_int]
    assert built_lib.b200_wan_block_workspace_bytes(75600, 5120, 13824) == 75600 * (5120 + 15360) * 2


def test_step_distill_scheduler_matches_reference_fixture(golden_dir):
    """host/wan_scheduler.py:WanStepDistillScheduler vs the REAL class (4-step sampler of BASELINE config 3): sigma / timestep grid, x0
    prediction and the re-noising from the default generator, bit for bit."""
    from safetensors import safe_open

    from lightx2v_b200.host.wan_scheduler import WanStepDistillScheduler

    with safe_open(os.path.join(golden_dir, "wan_scheduler_step_distill.safetensors"), framework="pt") as f:
        T = {k: f.get_tensor(k) for k in f.keys()}
        meta = f.metadata()
    cfg = dict(infer_steps=4, sample_shift=5.0, seed=42, target_shape=(16, 3, 8, 8), patch_size=(1, 2, 2), denoising_step_list=[1000, 750, 500, 250])
    sch = WanStepDistillScheduler(cfg, device="cpu")
    sch.prepare()
    assert torch.equal(sch.latents, T["latents_0"])
    assert torch.equal(sch.timesteps, T["timesteps"]) and torch.equal(sch.sigmas, T["sigmas"])
    for i in range(4):
        sch.step_pre(i)
        sch.noise_pred = T[f"noise_pred_{i}"]
        torch.manual_seed(int(meta["step_seed_base"]) + i)
        sch.step_post()
        assert sch.latents.dtype == T[f"latents_post_{i}"].dtype
        assert torch.equal(sch.latents, T[f"latents_post_{i}"]), i


def test_wan_model_reference_constructor_and_infer_class_selection(tmp_path):
    """WanModel(model_path, config, device) like the reference (wan/model.py:33-59): loads *.safetensors, picks the transformer infer class
    from feature_caching / model_cls (_init_infer_class, :61-75), and set_scheduler reaches the transformer infer (:180-185)."""
    from safetensors.torch import save_file

    from lightx2v_b200.host.wan_causvid import WanTransformerInferCausVid
    from lightx2v_b200.host.wan_infer import WanTransformerInfer
    from lightx2v_b200.host.wan_model import WanModel
    from lightx2v_b200.host.wan_teacache import WanTransformerInferTeaCaching

    dim, ffn = 256, 512
    W = O.synth_block_weights(1, dim, ffn, seed=3)
    W.update(O.synth_prepost_weights(dim, 16, "t2v", seed=4))
    save_file({k: v.contiguous() for k, v in W.items()}, str(tmp_path / "model.safetensors"))
    base = dict(task="t2v", num_layers=1, num_heads=2, dim=dim, ffn_dim=ffn, freq_dim=256, text_len=512, in_dim=16, out_dim=16, mm_config={},
                target_shape=(16, 1, 4, 4))
    m = WanModel(str(tmp_path), dict(base, feature_caching="NoCaching"), "cpu")
    assert type(m.transformer_infer) is WanTransformerInfer
    assert set(m.W.keys()) == set(W.keys()) and all(torch.equal(m.W[k], W[k]) for k in W)
    tea = WanModel(str(tmp_path), dict(base, feature_caching="Tea", teacache_thresh=0.2, coefficients=[[1.0, 0.0], [1.0, 0.0]], use_ret_steps=False,
                                       infer_steps=4, enable_cfg=True), "cpu")
    assert type(tea.transformer_infer) is WanTransformerInferTeaCaching
    cv = WanModel(str(tmp_path), dict(base, model_cls="wan2.1_causvid", num_frames=3, num_frame_per_block=1, frame_seq_length=4), "cpu")
    assert type(cv.transformer_infer) is WanTransformerInferCausVid
    with pytest.raises(NotImplementedError):
        WanModel(str(tmp_path), dict(base, feature_caching="TaylorSeer"), "cpu")
    with pytest.raises(FileNotFoundError):
        WanModel(str(tmp_path / "nope"), dict(base), "cpu")

    class Sched:
        infer_steps = 4

    s = Sched()
    for model in (m, tea, cv):
        model.set_scheduler(s)
        assert model.pre_infer.scheduler is s and model.post_infer.scheduler is s and model.transformer_infer.scheduler is s
    assert s.caching_records == [True] * 4                      # TeaCache installs its per-step records on the scheduler (schedulers/scheduler.py:11)


def test_weight_cache_is_rebuilt_when_the_weights_change():
    """derived tensors (concatenated QKV, native pointer struct, cached text K/V) are keyed on the phase object AND a fingerprint
    of its weight tensors, so a reload / replacement on the same tree cannot keep computing with the old tensors."""
    from lightx2v_b200.host.wan_infer import WanTransformerInfer
    from lightx2v_b200.host.wan_weights import WanTransformerWeights

    dim, ffn = 256, 512
    cfg = dict(task="t2v", num_layers=1, num_heads=2, dim=dim, ffn_dim=ffn, mm_config={})
    tree = WanTransformerWeights(cfg)
    tree.load(O.synth_block_weights(1, dim, ffn, seed=5))
    infer = WanTransformerInfer(cfg)
    sa = tree.blocks[0].compute_phases[1]
    c1 = infer._cache(sa)
    assert infer._cache(sa) is c1                              # stable while nothing changes
    tree.load(O.synth_block_weights(1, dim, ffn, seed=6))      # new tensors on the same tree
    c2 = infer._cache(sa)
    assert c2 is not c1 and c2.owner is sa
    sa.self_attn_q.bias.add_(1)                                # in-place edit bumps _version
    assert infer._cache(sa) is not c2
    infer.clear_weight_caches()
    assert not infer._caches


def test_install_into_the_real_reference_registries_and_uninstall():
    """registry.install_into_lightx2v() against the REAL LightX2V registries (vendored copy / build container; CPU shims of
    oracle/ref_loader.py): the B200 keys appear, the two hard-coded norm keys are overridden, the reference's own weight tree then
    builds B200 op objects from a stock config, and uninstall puts the reference's norm classes back."""
    from oracle import ref_loader

    if not ref_loader.import_reference():
        pytest.skip("reference package not available here")
    from lightx2v.utils import registry_factory as rf

    from lightx2v_b200.host import ops, registry

    ref_rms, ref_ln = rf.RMS_WEIGHT_REGISTER["sgl-kernel"], rf.LN_WEIGHT_REGISTER["Default"]
    assert ref_rms is not ops.RMSWeightB200 and ref_ln is not ops.LNWeightB200
    try:
        assert registry.install_into_lightx2v() is True
        assert rf.MM_WEIGHT_REGISTER[registry.MM_KEY] is ops.MMWeightB200 and rf.ATTN_WEIGHT_REGISTER[registry.ATTN_KEY] is ops.FmhaWeightB200
        assert rf.MM_WEIGHT_REGISTER[ops.FP8_MM_KEY] is ops.MMWeightFp8B200 and rf.MM_WEIGHT_REGISTER[ops.NVFP4_MM_KEY] is ops.MMWeightNvfp4B200
        assert rf.RMS_WEIGHT_REGISTER["sgl-kernel"] is ops.RMSWeightB200 and rf.LN_WEIGHT_REGISTER["Default"] is ops.LNWeightB200
        assert registry.install_into_lightx2v() is True                      # idempotent: no "already exists"
        # the reference's OWN weight tree, driven by a stock config that names the B200 keys
        from lightx2v.models.networks.wan.weights.transformer_weights import WanTransformerWeights

        cfg = ref_loader.ref_config(dim=128, num_heads=1, ffn_dim=256, num_layers=1, mm_type=registry.MM_KEY, attn_type=registry.ATTN_KEY)
        tree = WanTransformerWeights(cfg)
        kinds = {type(m).__name__ for blk in tree.blocks for phase in blk.compute_phases for m in phase._modules.values()}
        assert {"MMWeightB200", "FmhaWeightB200", "RMSWeightB200"} <= kinds, kinds
    finally:
        registry.uninstall_from_lightx2v()
    assert rf.RMS_WEIGHT_REGISTER["sgl-kernel"] is ref_rms and rf.LN_WEIGHT_REGISTER["Default"] is ref_ln
