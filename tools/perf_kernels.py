#!/usr/bin/env python
"""Kernel-level timing sweeps (one process per env setting). Usage: python tools/perf_kernels.py fmha|gemm"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def child(kind):
    import torch
    from lightx2v_b200 import lib
    from tools.bringup import _time_cuda
    if kind in ("fmha", "fmha3", "fmha4", "fmha5", "fmha6", "fmha7"):
        for (S, H) in ((75600, 40),):
            q = torch.randn(S, H, 128, device="cuda").bfloat16(); k = torch.randn(S, H, 128, device="cuda").bfloat16(); v = torch.randn(S, H, 128, device="cuda").bfloat16()
            o = torch.empty_like(q)
            ms = _time_cuda(lambda: lib.fmha(q, k, v, out=o), iters=3, warmup=1)
            print(json.dumps({"case": f"fmha_S{S}_H{H}", "env": os.environ.get("B200_FMHA_POLY"), "ver": os.environ.get("B200_FMHA_VER"), "ms": ms, "tflops": 4 * S * S * H * 128 / ms / 1e9}))
            ref = torch.nn.functional.scaled_dot_product_attention(q[:2048].float().transpose(0, 1), k[:4096].float().transpose(0, 1), v[:4096].float().transpose(0, 1)).transpose(0, 1)
            got = lib.fmha(q[:2048], k[:4096], v[:4096])
            print(json.dumps({"case": "fmha_err", "env": os.environ.get("B200_FMHA_POLY"), "max_abs_err": float((got.float() - ref).abs().max())}))
    elif kind == "gemm4":
        for (M, N, K) in ((75600, 5120, 5120), (75600, 13824, 5120), (75600, 5120, 13824), (75600, 15360, 5120)):
            a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16(); b = torch.randn(N, device="cuda").bfloat16()
            o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            gw, _ = lib.nvfp4_act_scale(w)
            wq, sw = lib.quant_nvfp4(w, gw)
            ga, alpha = lib.nvfp4_act_scale(a, gw)
            aq, sa = lib.quant_nvfp4(a, ga)
            for bn in (128, 256):
                ms = _time_cuda(lambda: lib.gemm_nvfp4(aq, wq, sa, sw, alpha, b, out=o, block_n=bn))
                print(json.dumps({"case": f"gemm_nvfp4_{M}x{N}x{K}_bn{bn}", "ms": ms, "tflops": 2 * M * N * K / ms / 1e9}))
            ms = _time_cuda(lambda: lib.quant_nvfp4(a, ga))
            print(json.dumps({"case": f"quant_nvfp4_{M}x{K}", "ms": ms, "gbs": M * K * (2 + 0.5 + 1 / 16) / ms / 1e6}))
            ms = _time_cuda(lambda: lib.nvfp4_act_scale(a, gw))
            print(json.dumps({"case": f"nvfp4_act_scale_{M}x{K}", "ms": ms, "gbs": M * K * 2 / ms / 1e6}))
    elif kind == "gemm8":
        for (M, N, K) in ((75600, 5120, 5120), (75600, 13824, 5120), (75600, 5120, 13824), (75600, 15360, 5120)):
            a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16(); b = torch.randn(N, device="cuda").bfloat16()
            o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            aq, sa = lib.quant_fp8_per_token(a)
            wq, sw = lib.quant_fp8_per_token(w)
            ms = _time_cuda(lambda: lib.gemm_fp8(aq, sa, wq, sw, b, out=o))
            print(json.dumps({"case": f"gemm_fp8_{M}x{N}x{K}", "ms": ms, "tflops": 2 * M * N * K / ms / 1e9}))
            ms = _time_cuda(lambda: lib.quant_fp8_per_token(a, out=aq, scale=sa))
            print(json.dumps({"case": f"quant_fp8_{M}x{K}", "ms": ms, "gbs": M * K * 3 / ms / 1e6}))
    else:
        for (M, N, K) in ((75600, 5120, 5120), (75600, 13824, 5120), (75600, 5120, 13824), (75600, 15360, 5120)):
            a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16(); b = torch.randn(N, device="cuda").bfloat16()
            o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            ms = _time_cuda(lambda: lib.gemm_bf16(a, w, b, out=o))
            print(json.dumps({"case": f"gemm_{M}x{N}x{K}", "env": os.environ.get("B200_GEMM_GROUP_M"), "ms": ms, "tflops": 2 * M * N * K / ms / 1e9}))

if __name__ == "__main__":
    if len(sys.argv) > 2:
        child(sys.argv[1])
    else:
        kind = sys.argv[1]
        var, vals = {"fmha": ("B200_FMHA_POLY", ["0", "1", "2"]), "fmha3": ("B200_FMHA_POLY", ["0", "1", "2"]), "fmha4": ("B200_FMHA_POLY", ["1"]), "fmha5": ("B200_FMHA_POLY", ["0", "1", "2"]), "fmha6": ("B200_FMHA_POLY", ["0", "1", "2"]), "fmha7": ("B200_FMHA_POLY", ["0", "1"]), "gemm": ("B200_GEMM_GROUP_M", ["16"]), "gemm8": ("B200_X", ["0"]), "gemm4": ("B200_X", ["0"])}[kind]
        for v in vals:
            env = dict(os.environ); env[var] = v
            if kind == "fmha": env["B200_FMHA_VER"] = "2"
            if kind == "fmha3": env["B200_FMHA_VER"] = "3"
            if kind == "fmha4": env["B200_FMHA_VER"] = "4"
            if kind == "fmha5": env["B200_FMHA_VER"] = "5"
            if kind == "fmha6": env["B200_FMHA_VER"] = "6"
            if kind == "fmha7": env["B200_FMHA_VER"] = "7"
            subprocess.run([sys.executable, os.path.abspath(__file__), kind, "child"], env=env, cwd=ROOT)
