"""CUDA-event timing of the bf16 / fp8 GEMM at the 14B block's four shapes (and cuBLAS bf16 beside it)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lightx2v_b200 import lib  # noqa: E402


def t(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


for (M, N, K) in ((75600, 5120, 5120), (75600, 13824, 5120), (75600, 5120, 13824), (75600, 15360, 5120)):
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16()
    o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ms = t(lambda: lib.gemm_bf16(a, w, b, out=o))
    ms_cublas = t(lambda: torch.addmm(b, a, w.t(), out=o))
    aq, sa = lib.quant_fp8_per_token(a)
    wq, sw = lib.quant_fp8_per_token(w)
    ms8 = t(lambda: lib.gemm_fp8(aq, sa, wq, sw, b, out=o))
    fl = 2.0 * M * N * K / 1e9
    print(f"{M}x{N}x{K}: bf16 {ms:.3f} ms {fl / ms:.0f} TFLOP/s | cuBLAS {ms_cublas:.3f} ms {fl / ms_cublas:.0f} | fp8 {ms8:.3f} ms {fl / ms8:.0f}", flush=True)
    del a, w, o, aq, wq
