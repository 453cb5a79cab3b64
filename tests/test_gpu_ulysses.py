"""2-GPU NCCL test of the Ulysses path (skipped on a 1-GPU box): the sharded fused block stack must reproduce the
single-GPU result on the same inputs, including a token count that is not a multiple of the world size (padded keys masked)."""
import os
import socket

import pytest
import torch

from oracle import wan_oracle as O

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_rank(rank, world, port, grid, out_path, fused=False):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from lightx2v_b200 import lib
        from lightx2v_b200.host import ulysses as U
        from lightx2v_b200.host.wan_infer import WanTransformerInfer
        from lightx2v_b200.host.wan_weights import WanTransformerWeights

        dim, heads, ffn, L = 1536, 12, 8960, 2
        S = grid[0] * grid[1] * grid[2]
        W = O.synth_block_weights(L, dim, ffn, seed=1, device="cuda")
        x, embed0, context = O.synth_block_inputs(S, dim, seed=2, device="cuda")
        cfg = dict(task="t2v", num_layers=L, num_heads=heads, dim=dim, ffn_dim=ffn, mm_config={})
        weights = WanTransformerWeights(cfg)
        weights.load(W)
        freqs = O.wan_freqs_table(128)
        g = torch.tensor([grid])
        ref = WanTransformerInfer(cfg).infer(weights, g, None, x.clone(), embed0, None, freqs, context)       # single GPU
        infer = WanTransformerInfer(cfg)
        if fused:
            infer.parallel_attention = U.FusedUlyssesAttention(S, U.shard_rows(S, world), heads, torch.device("cuda", rank))
        else:
            infer.parallel_attention = U.UlyssesAttention(lib.fmha, total_rows=S)
        infer.sp_rank, infer.sp_world = rank, world
        xs = U.pre_process(x.clone(), rank, world)
        ys = infer.infer(weights, g, None, xs, embed0, None, freqs, context)
        full = U.post_process(ys, S)
        torch.cuda.synchronize()
        if rank == 0:
            err = (full.float() - ref.float()).abs()
            bad = (err > 1e-2 + 1e-2 * ref.float().abs()).float().mean().item()
            torch.save({"bad": bad, "max": err.max().item()}, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("fused", [False, True])               # NCCL all-to-all path / peer-memory kernels
@pytest.mark.parametrize("grid", [(4, 8, 10), (3, 7, 9)])        # 320 tokens (even) and 189 tokens (padded to 190)
def test_ulysses_blocks_match_single_gpu(tmp_path, grid, fused):
    import torch.multiprocessing as mp

    out = str(tmp_path / "res.pt")
    mp.spawn(_run_rank, args=(2, _free_port(), grid, out, fused), nprocs=2, join=True)
    r = torch.load(out)
    print("ulysses vs single GPU:", r)
    assert r["bad"] < 2e-3 and r["max"] < 0.13
