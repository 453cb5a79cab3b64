#!/usr/bin/env python
"""CUDA-event timings of the HBM-bound (streaming) kernels at their bench shapes -> achieved GB/s and fraction of the measured copy peak.
Inputs are far larger than the 126 MB L2, so every iteration streams from HBM.  One JSON line per kernel on stdout."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lightx2v_b200 import lib  # noqa: E402


def peak():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    lib.load()
    pk = peak()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g, device=dev) * scale).to(torch.bfloat16)

    out = []

    def rep(name, nbytes, ms, note=""):
        gbs = nbytes / (ms * 1e-3) / 1e9
        r = {"kernel": name, "algorithmic_bytes": nbytes, "ms": round(ms, 4), "GBps": round(gbs, 1), "frac_of_measured_hbm_peak": round(gbs / pk, 3), "note": note}
        out.append(r)
        print(json.dumps(r), flush=True)

    # ---- Wan VAE: rms_silu at the three decoder stages (frames x H x W x C of the 81 x 720 x 1280 decode)
    for C, T, H, W in ((384, 21, 90, 160), (384, 41, 180, 320), (192, 81, 360, 640), (96, 81, 720, 1280)):
        vox = T * H * W
        if vox * C * 2 > 8e9:
            T = max(1, int(8e9 / (H * W * C * 2)))
            vox = T * H * W
        x = rnd(T, H, W, C)
        y = torch.empty_like(x)
        gm = torch.ones(C, device=dev)
        ms = timeit(lambda: lib.rms_silu_cl(x, gm, out=y))
        rep(f"rms_silu_kernel<{C}>", vox * C * 4, ms, f"[{T},{H},{W},{C}]")
        del x, y
    # ---- Hunyuan VAE: gn_stats + gn_apply_pad on one full-width tile stage
    for C, T, H, W in ((512, 17, 32, 32), (256, 65, 128, 128), (128, 65, 256, 256)):
        x = rnd(T, H, W, C)
        sums = lib.gn_stats_cl(x)
        gm, bt = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        yp = torch.empty(T + 2, H + 2, W + 2, C, dtype=torch.bfloat16, device=dev)
        ms = timeit(lambda: lib.gn_stats_cl(x, sums))
        rep(f"gn_stats_kernel<{C}>", x.numel() * 2, ms, f"[{T},{H},{W},{C}]")
        ms = timeit(lambda: lib.gn_apply_pad_cl(x, sums, gm, bt, pad=(2, 1, 1), out=yp))
        rep(f"gn_apply_pad_kernel<{C}>", x.numel() * 2 + yp.numel() * 2, ms, f"[{T},{H},{W},{C}] -> padded")
        del x, yp
    # ---- DiT row-wise kernels at S = 75 600, D = 5120
    S, D = 75600, 5120
    x = rnd(S, D)
    sc, sh = rnd(D, scale=0.1), rnd(D, scale=0.1)
    y = torch.empty_like(x)
    ms = timeit(lambda: lib.ln_modulate(x, scale=sc, shift=sh, out=y))
    rep("ln_modulate_kernel", S * D * 4, ms)
    q8 = torch.empty(S, D, dtype=lib.FP8, device=dev)
    s8 = torch.empty(S, 1, dtype=torch.float32, device=dev)
    ms = timeit(lambda: lib.quant_fp8_per_token(x, out=q8, scale=s8))
    rep("quant_fp8_kernel", S * D * 3, ms)
    ms = timeit(lambda: lib.ln_modulate_fp8(x, scale=sc, shift=sh, out=q8, out_scale=s8))
    rep("ln_modulate_kernel<fp8 out>", S * D * 3, ms)
    del y, q8
    qkv = rnd(S, 3 * D)
    wq, wk = 1 + rnd(D, scale=0.05), 1 + rnd(D, scale=0.05)
    cs = torch.randn(S, 64, 2, device=dev)
    ms = timeit(lambda: lib.rms_rope_(qkv[:, :D], wq, qkv[:, D:2 * D], wk, cos_sin=cs))
    rep("rms_rope_kernel<rope>", S * D * 2 * 4 + cs.numel() * 4, ms, "q and k in place inside the fused qkv buffer")
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "perf_stream.jsonl")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        for r in out:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
