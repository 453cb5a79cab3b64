#!/usr/bin/env python
"""Time the HunyuanVideo VAE decode (720p x 129 frames: latent [1,16,33,90,160], 3 temporal x 28 spatial tiles) on one B200, and one
full tile [16,17,32,32] against the torch restatement of the reference (fp16, cuDNN) on the same GPU.  Usage: python tools/perf_hunyuan_vae.py [--small]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hunyuan_vae_oracle as HV  # noqa: E402  (tools/ are measurement scripts, not the product path)
from lightx2v_b200.host.hunyuan_vae import HunyuanVAEB200  # noqa: E402


def ev_time(fn, iters=1, warmup=1):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    small = "--small" in sys.argv
    cfg = dict(HV.HUNYUAN_VAE_CFG)
    W = {k: v.cuda() for k, v in HV.synth_vae_weights(cfg, seed=5).items()}
    vae = HunyuanVAEB200(W, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    tile = torch.randn(1, 16, 17, 32, 32, generator=g, device="cuda")
    ms_tile = ev_time(lambda: vae.decoder.decode_tile(tile[0]), iters=2)
    print(json.dumps({"case": "hunyuan_vae_tile_17x32x32", "ms": ms_tile, "out": [3, 65, 256, 256]}), flush=True)
    Wh = {k: v.half() for k, v in W.items()}
    with torch.no_grad():
        zt = (tile / cfg["scaling_factor"]).half()
        ms_ref = ev_time(lambda: HV.tile_decode(Wh, zt, cfg), iters=1)
        ref = HV.tile_decode(Wh, zt, cfg)[0].float()
    out = vae.decoder.decode_tile(tile[0])
    mse = (out - ref).pow(2).mean()
    print(json.dumps({"case": "reference_port_fp16_tile_17x32x32", "ms": ms_ref, "speedup": ms_ref / ms_tile,
                      "psnr_vs_fp16_port_db": float(10 * torch.log10(ref.abs().max() ** 2 / mse))}), flush=True)
    if small:
        return
    lat = torch.randn(1, 16, 33, 90, 160, generator=g, device="cuda")
    torch.cuda.reset_peak_memory_stats()
    ms = ev_time(lambda: vae.decode_device(lat), iters=1, warmup=1)
    t0 = time.time()
    img = vae.decode(lat)
    wall = time.time() - t0
    mpix = img.shape[2] * img.shape[3] * img.shape[4] / 1e6
    print(json.dumps({"case": "hunyuan_vae_decode_720p_129f", "ms_device": ms, "mpix": mpix, "mpix_per_s": mpix / (ms * 1e-3), "wall_s_with_d2h": wall,
                      "out": list(img.shape), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                      "reference_port_estimate_s": (ms_ref / ms_tile) * ms * 1e-3}), flush=True)


if __name__ == "__main__":
    main()
