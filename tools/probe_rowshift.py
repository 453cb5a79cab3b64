"""UMMA row-shift probe (csrc/probe.cu): which descriptor addressing does the tensor core honour for a K-major swizzled operand that
starts r0 rows into a TMA-written tile?  Prints max |D - ref| for r0 = 0..8, both swizzle widths, base_offset = 0 vs r0 % 8."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lightx2v_b200 import lib  # noqa: E402

L = lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
for k in (64, 32):
    A = torch.randn(136, k, generator=g, device="cuda").to(torch.bfloat16)
    B = torch.randn(64, k, generator=g, device="cuda").to(torch.bfloat16)
    for mode in (0, 1):
        errs = []
        for r0 in range(9):
            D = torch.zeros(128, 64, dtype=torch.float32, device="cuda")
            rc = L.b200_debug_umma_rowshift(A.data_ptr(), B.data_ptr(), D.data_ptr(), k, r0, mode, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            ref = A[r0:r0 + 128].float() @ B.float().t()
            errs.append(round((D - ref).abs().max().item(), 4) if rc == 0 else f"rc={rc}")
        print(f"k={k} ({k * 2}-byte rows) base_offset={'r0%8' if mode else '0'}: max|err| for r0=0..8:", errs, flush=True)
