"""UMMA issue-rate probe (csrc/probe.cu): clocks per SS-mode tcgen05.mma (M = 128, K = 16, bf16) at each N with operands resident in shared
memory - the floor the conv / GEMM tile shapes are judged against.  One CTA per SM on `grid` SMs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lightx2v_b200 import lib  # noqa: E402

L = lib.load()
ITERS = 2000


def run(n, n_acc=1, issuers=1, writers=0, a_tiles=1, grid=148):
    out = torch.zeros(3 * grid, dtype=torch.int64, device="cuda")
    rc = L.b200_debug_umma_rate(n, ITERS, n_acc, issuers, writers, a_tiles, grid, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    if rc:
        return None
    torch.cuda.synchronize()
    o = out.cpu()
    clk = o[: 2 * grid].view(grid, 2)[:, :issuers].float().max(dim=1).values.median().item()
    per_mma = clk / (ITERS * 4 * issuers)
    st = o[2 * grid:].float().median().item() * 512 / max(clk, 1)
    return per_mma, st


print(f"{'N':>4} {'acc':>3} {'iss':>3} {'wr':>2} {'At':>2} | clk/MMA  floor(N/2)  pipe%   smem rd B/clk   writer st B/clk", flush=True)
for grid in (148, 1):
    print(f"-- grid {grid}")
    for n in (16, 32, 64, 96, 128, 192, 256):
        for (n_acc, issuers, writers, a_tiles) in ((1, 1, 0, 1), (2, 1, 0, 1), (1, 2, 0, 1), (2, 1, 0, 4), (2, 1, 2, 4), (2, 1, 4, 4)):
            if issuers * n_acc * n > 512:
                continue
            if grid == 1 and (writers or a_tiles > 1):
                continue
            r = run(n, n_acc, issuers, writers, a_tiles, grid)
            if r is None:
                print(n, n_acc, issuers, writers, a_tiles, "rc!=0", L.b200_last_error())
                continue
            per, st = r
            rd = (4096 + n * 32) / per
            print(f"{n:>4} {n_acc:>3} {issuers:>3} {writers:>2} {a_tiles:>2} | {per:7.1f}  {n / 2:9.1f}  {100 * (n / 2) / per:5.1f}   {rd:8.1f}        {st:8.1f}", flush=True)
