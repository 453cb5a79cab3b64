"""Wan VAE decode of the bench latent for several chunk sizes: time, MPix/s, peak memory."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lightx2v_b200.host.wan_vae import WanVAEDecoderB200  # noqa: E402
from oracle import vae_oracle as V   # synthetic weights only  # noqa: E402

W = V.synth_vae_weights(0)
zs = torch.randn(16, 21, 90, 160, device="cuda")
mpix = 81 * 720 * 1280 / 1e6
for chunk in [int(v) for v in sys.argv[1:]] or [2, 3, 5, 7]:
    dec = WanVAEDecoderB200(W, device="cuda", chunk_frames=chunk if chunk > 0 else None)
    dec.decode(zs)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(2):
        out = dec.decode(zs)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 2
    print(json.dumps({"chunk_frames": chunk, "ms": round(ms, 1), "MPix/s": round(mpix / ms * 1e3, 1), "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 1)}), flush=True)
    del dec, out
    torch.cuda.empty_cache()
