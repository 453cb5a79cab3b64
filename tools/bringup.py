#!/usr/bin/env python
"""GPU bring-up harness: runs each kernel check in its own subprocess (a trap or hang in one case cannot take the
others down), with a per-case timeout, and writes gpurun_out/bringup.json + per-case logs.

    python tools/bringup.py            # driver: all cases
    python tools/bringup.py --case X   # one case in-process
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def _report_err(name, got, ref, atol, rtol):
    import torch
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    res = {
        "case": name,
        "max_abs_err": float(err.max()),
        "mean_abs_err": float(err.mean()),
        "ref_absmax": float(ref.abs().max()),
        "bad_frac": float(bad.float().mean()),
        "nan": int(torch.isnan(got).sum()),
        "ok": bool(not bad.any()) and not bool(torch.isnan(got).any()),
    }
    if bad.any() and got.dim() == 2:
        idx = bad.nonzero()[:8].tolist()
        res["first_bad"] = [(i, j, float(got[i, j]), float(ref[i, j])) for i, j in idx]
        # error maps to localise layout bugs: per 8-row group and per 16-col group
        R, C = got.shape
        rb = bad.float().view(R // 8 if R % 8 == 0 else 1, -1).mean(1) if R % 8 == 0 else None
        if rb is not None:
            res["bad_by_rowgroup8_head"] = [round(float(x), 3) for x in rb[:16]]
        if C % 16 == 0:
            cb = bad.float().view(R, C // 16, 16).mean((0, 2))
            res["bad_by_colgroup16_head"] = [round(float(x), 3) for x in cb[:16]]
    print(json.dumps(res))
    return res


def case_rowwise():
    import torch
    from lightx2v_b200 import lib
    torch.manual_seed(0)
    out = []
    for rows, D in ((300, 1536), (257, 5120), (64, 3072)):
        x = torch.randn(rows, D, device="cuda").bfloat16() * 2 + 0.3
        scale = (torch.randn(D, device="cuda") * 0.1).bfloat16()
        shift = (torch.randn(D, device="cuda") * 0.1).bfloat16()
        w = (1 + torch.randn(D, device="cuda") * 0.1).bfloat16()
        b = (torch.randn(D, device="cuda") * 0.1).bfloat16()
        # modulated, no affine
        y = lib.ln_modulate(x, scale=scale, shift=shift)
        ref = torch.nn.functional.layer_norm(x, (D,), None, None, 1e-6)
        ref = ref.mul_(1 + scale).add_(shift)
        out.append(_report_err(f"ln_mod_{rows}x{D}", y, ref, 2e-2, 2e-2))
        y = lib.ln_modulate(x, weight=w, bias=b)
        ref = torch.nn.functional.layer_norm(x, (D,), w, b, 1e-6)
        out.append(_report_err(f"ln_affine_{rows}x{D}", y, ref, 2e-2, 2e-2))
        # rms + rope
        if D % 128 == 0:
            q = torch.randn(rows, D, device="cuda").bfloat16()
            k = torch.randn(rows, D, device="cuda").bfloat16()
            wq = (1 + torch.randn(D, device="cuda") * 0.1).bfloat16()
            wk = (1 + torch.randn(D, device="cuda") * 0.1).bfloat16()
            ang = torch.rand(rows, 64, device="cuda", dtype=torch.float64) * 6.28
            cs = torch.stack([ang.cos(), ang.sin()], -1).float().contiguous()

            def ref_rms(t, w):
                t = t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + 1e-6)
                return t * w

            def ref_rope(t):
                H = D // 128
                tc = torch.view_as_complex(t.to(torch.float64).reshape(rows, H, 64, 2))
                f = torch.polar(torch.ones_like(ang), ang).view(rows, 1, 64)
                return torch.view_as_real(tc * f).flatten(2).to(torch.bfloat16).reshape(rows, D)

            rq, rk = ref_rope(ref_rms(q, wq)), ref_rope(ref_rms(k, wk))
            lib.rms_rope_(q, wq, k, wk, cos_sin=cs)
            out.append(_report_err(f"rms_rope_q_{rows}x{D}", q, rq, 2e-2, 2e-2))
            out.append(_report_err(f"rms_rope_k_{rows}x{D}", k, rk, 2e-2, 2e-2))
    return out


def _gemm_case(M, N, K, epi, block_n, seed=0):
    import torch
    from lightx2v_b200 import lib
    torch.manual_seed(seed)
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda").bfloat16()
    gate = torch.randn(N, device="cuda").bfloat16()
    x = torch.randn(M, N, device="cuda").bfloat16()
    y = (a.float() @ w.float().t() + bias.float()).bfloat16()
    if epi == lib.EPI_BIAS:
        out = lib.gemm_bf16(a, w, bias, epilogue=epi, block_n=block_n)
        ref = y
    elif epi == lib.EPI_BIAS_GELU:
        out = lib.gemm_bf16(a, w, bias, epilogue=epi, block_n=block_n)
        ref = torch.nn.functional.gelu(y, approximate="tanh")
    elif epi == lib.EPI_GATE_RESIDUAL:
        out = x.clone()
        lib.gemm_bf16(a, w, bias, out=out, epilogue=epi, gate=gate, block_n=block_n)
        ref = x + y * gate
    else:
        out = x.clone()
        lib.gemm_bf16(a, w, bias, out=out, epilogue=epi, block_n=block_n)
        ref = x + y
    torch.cuda.synchronize()
    return _report_err(f"gemm_M{M}_N{N}_K{K}_epi{epi}_bn{block_n}", out, ref, 3e-2, 2e-2)


def case_gemm_small():
    out = []
    for bn in (128, 256):
        out.append(_gemm_case(128, 256, 64, 0, bn))
        out.append(_gemm_case(256, 512, 256, 0, bn))
    return out


def case_gemm_shapes():
    out = []
    for (M, N, K) in ((4096 + 80, 5120, 5120), (1000, 1536, 1536), (777, 8960, 1536), (512, 5120, 13824), (300, 64, 1536)):
        for epi in (0, 1, 2, 3):
            try:
                out.append(_gemm_case(M, N, K, epi, 0))
            except Exception as ex:
                r = {"case": f"gemm_M{M}_N{N}_K{K}_epi{epi}", "ok": False, "exception": str(ex)[:300]}
                print(json.dumps(r)); out.append(r)
                if "CUDA" in str(ex) or "cuda" in str(ex):
                    return out
    return out


def _fmha_case(sq, sk, H, seed=0, strided=False):
    import torch
    from lightx2v_b200 import lib
    torch.manual_seed(seed)
    if strided:
        qkv = torch.randn(max(sq, sk), 3, H, 128, device="cuda").bfloat16()
        q, k, v = qkv[:sq, 0], qkv[:sk, 1], qkv[:sk, 2]
    else:
        q = torch.randn(sq, H, 128, device="cuda").bfloat16()
        k = torch.randn(sk, H, 128, device="cuda").bfloat16()
        v = torch.randn(sk, H, 128, device="cuda").bfloat16()
    out = lib.fmha(q, k, v)
    torch.cuda.synchronize()
    ref = torch.nn.functional.scaled_dot_product_attention(
        q.float().transpose(0, 1), k.float().transpose(0, 1), v.float().transpose(0, 1)).transpose(0, 1)
    return _report_err(f"fmha_sq{sq}_sk{sk}_H{H}{'_strided' if strided else ''}", out.reshape(sq, -1), ref.reshape(sq, -1), 2e-2, 2e-2)


def case_fmha_small():
    return [_fmha_case(256, 128, 1), _fmha_case(256, 256, 1), _fmha_case(256, 512, 2)]


def case_fmha_shapes():
    out = [_fmha_case(1000, 1000, 2), _fmha_case(4096, 4096, 12, strided=True), _fmha_case(3000, 512, 4), _fmha_case(700, 257, 3),
           _fmha_case(130, 77, 1)]
    return out


def _time_cuda(fn, iters=5, warmup=2):
    import torch
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


def case_perf_gemm():
    import torch
    from lightx2v_b200 import lib
    out = []
    for (M, N, K) in ((75600, 5120, 5120), (75600, 13824, 5120), (75600, 5120, 13824), (75600, 15360, 5120)):
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        bias = torch.randn(N, device="cuda").bfloat16()
        o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for bn in (256, 128):
            ms = _time_cuda(lambda: lib.gemm_bf16(a, w, bias, out=o, block_n=bn))
            r = {"case": f"perf_gemm_{M}x{N}x{K}_bn{bn}", "ms": ms, "tflops": 2 * M * N * K / ms / 1e9, "ok": True}
            print(json.dumps(r)); out.append(r)
        ms = _time_cuda(lambda: torch.addmm(bias, a, w.t(), out=o))
        r = {"case": f"perf_cublas_{M}x{N}x{K}", "ms": ms, "tflops": 2 * M * N * K / ms / 1e9, "ok": True}
        print(json.dumps(r)); out.append(r)
    return out


def case_perf_fmha():
    import torch
    from lightx2v_b200 import lib
    out = []
    for (S, H) in ((32760, 12), (75600, 40)):
        q = torch.randn(S, H, 128, device="cuda").bfloat16()
        k = torch.randn(S, H, 128, device="cuda").bfloat16()
        v = torch.randn(S, H, 128, device="cuda").bfloat16()
        o = torch.empty_like(q)
        ms = _time_cuda(lambda: lib.fmha(q, k, v, out=o), iters=2, warmup=1)
        r = {"case": f"perf_fmha_S{S}_H{H}", "ms": ms, "tflops": 4 * S * S * H * 128 / ms / 1e9, "ok": True}
        print(json.dumps(r)); out.append(r)
        try:
            from flash_attn import flash_attn_varlen_func
            cu = torch.tensor([0, S], dtype=torch.int32, device="cuda")
            ms = _time_cuda(lambda: flash_attn_varlen_func(q, k, v, cu, cu, S, S), iters=2, warmup=1)
            r = {"case": f"perf_fa2_S{S}_H{H}", "ms": ms, "tflops": 4 * S * S * H * 128 / ms / 1e9, "ok": True}
            print(json.dumps(r)); out.append(r)
            ref = flash_attn_varlen_func(q, k, v, cu, cu, S, S)
            out.append(_report_err(f"fmha_vs_fa2_S{S}_H{H}", o.reshape(S, -1), ref.reshape(S, -1), 1e-2, 1e-2))
        except Exception as ex:  # noqa
            print("flash_attn unavailable:", ex)
    return out


CASES = {
    "rowwise": (case_rowwise, 120),
    "gemm_small": (case_gemm_small, 120),
    "gemm_shapes": (case_gemm_shapes, 240),
    "fmha_small": (case_fmha_small, 120),
    "fmha_shapes": (case_fmha_shapes, 240),
    "perf_gemm": (case_perf_gemm, 240),
    "perf_fmha": (case_perf_fmha, 300),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    if args.case:
        import traceback
        try:
            res = CASES[args.case][0]()
        except Exception:
            traceback.print_exc()
            sys.exit(2)
        bad = [r for r in res if not r.get("ok")]
        sys.exit(1 if bad else 0)
    os.makedirs(OUT, exist_ok=True)
    summary = {}
    names = [n for n in CASES if not args.only or n in args.only.split(",")]
    for name in names:
        t0 = time.time()
        log = os.path.join(OUT, f"bringup_{name}.log")
        with open(log, "w") as f:
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name], stdout=f,
                                   stderr=subprocess.STDOUT, timeout=CASES[name][1], cwd=ROOT)
                rc = p.returncode
            except subprocess.TimeoutExpired:
                rc = "timeout"
        lines = [json.loads(l) for l in open(log) if l.startswith("{")]
        summary[name] = {"rc": rc, "secs": round(time.time() - t0, 1), "results": lines}
        print(name, rc, f"{time.time() - t0:.1f}s", "ok" if rc == 0 else "FAIL")
        if rc != 0:
            tail = open(log).read()[-3000:]
            print(tail)
    json.dump(summary, open(os.path.join(OUT, "bringup.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
