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


def _run_hunyuan_rank(rank, world, port, out_path):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle import hunyuan_oracle as HO
        from lightx2v_b200 import lib
        from lightx2v_b200.host import ulysses as U
        from lightx2v_b200.host.hunyuan_infer import HunyuanTransformerInfer, HunyuanTransformerWeights

        hidden, mlp, heads = 3072, 12288, 24
        Li, Lt, valid = 1200, 256, 77
        W = HO.synth_weights(1, 1, hidden, mlp, seed=3, device="cuda")
        img, txt, vec, cu, freqs = HO.synth_inputs(Li, Lt, valid, hidden, seed=4, device="cuda")
        cfg = dict(task="t2v", mm_config={}, double_blocks_num=1, single_blocks_num=1)
        weights = HunyuanTransformerWeights(cfg)
        weights.load(W)
        cu_t = torch.tensor(cu, dtype=torch.int32)
        ref, _ = HunyuanTransformerInfer(cfg).infer(weights, img.clone(), txt.clone(), vec, cu_t, Li + Lt, freqs)      # single GPU
        infer = U.parallelize_hunyuan(HunyuanTransformerInfer(cfg), lib.fmha)
        s = Li // world
        sl = slice(rank * s, (rank + 1) * s)
        out, _ = infer.infer(weights, img[sl].clone(), txt.clone(), vec, cu_t, Li + Lt, (freqs[0][sl].contiguous(), freqs[1][sl].contiguous()))
        full = torch.empty(Li, hidden, dtype=out.dtype, device=out.device)
        dist.all_gather_into_tensor(full, out.contiguous())
        torch.cuda.synchronize()
        if rank == 0:
            err = (full.float() - ref.float()).abs()
            bad = (err > 1e-2 + 1e-2 * ref.float().abs()).float().mean().item()
            torch.save({"bad": bad, "max": err.max().item()}, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_hunyuan_ulysses_blocks_match_single_gpu(tmp_path):
    """HunyuanVideo double + single block with image tokens sharded over 2 GPUs and replicated text (two varlen segments)."""
    import torch.multiprocessing as mp

    out = str(tmp_path / "res_h.pt")
    mp.spawn(_run_hunyuan_rank, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    print("hunyuan ulysses vs single GPU:", r)
    assert r["bad"] < 2e-3 and r["max"] < 0.13


def _run_vae_rank(rank, world, port, out_path):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle import hunyuan_vae_oracle as HV
        from lightx2v_b200.host.hunyuan_vae import HunyuanVAEB200

        cfg = dict(HV.HUNYUAN_VAE_CFG, block_out_channels=(64, 64, 128, 128), sample_size=64, sample_tsize=16)
        vae = HunyuanVAEB200(HV.synth_vae_weights(cfg, seed=3), device="cuda", config=cfg)
        g = torch.Generator().manual_seed(11)
        lat = torch.randn(1, 16, 6, 12, 10, generator=g).cuda()
        single = vae.decode(lat)
        par = vae.decode_dist(lat)
        if rank == 0:
            torch.save({"equal": bool(torch.equal(single, par)), "max": float((single - par).abs().max())}, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_hunyuan_vae_tile_parallel_decode_is_bit_identical(tmp_path):
    """8 tiles (2 temporal x 2 x 2 spatial) decoded by 2 ranks alternately and broadcast == the single-GPU tiled decode, bit for bit."""
    import torch.multiprocessing as mp

    out = str(tmp_path / "res_v.pt")
    mp.spawn(_run_vae_rank, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    print("hunyuan VAE tile-parallel vs single GPU:", r)
    assert r["equal"], r


def _run_cfg_rank(rank, world, port, out_path):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import bench as B
        from lightx2v_b200 import lib
        from lightx2v_b200.host import ulysses as U
        from lightx2v_b200.host.wan_model import WanModel
        from lightx2v_b200.host.wan_scheduler import WanScheduler

        cfg = dict(dim=1536, num_heads=12, ffn_dim=8960, num_layers=2, target_shape=(16, 3, 16, 16), infer_steps=4, enable_cfg=True, sample_guide_scale=5.0,
                   sample_shift=5.0, task="t2v", freq_dim=256, text_len=512, in_dim=16, out_dim=16, seed=42, mm_config={}, patch_size=(1, 2, 2))
        dev = torch.device("cuda", rank)
        W = B.synth_weights(cfg, dev)
        g = torch.Generator(device=dev).manual_seed(7)
        ctx = {"context": torch.randn(512, 4096, generator=g, device=dev).to(torch.bfloat16), "context_null": torch.randn(512, 4096, generator=g, device=dev).to(torch.bfloat16)}
        inputs = {"text_encoder_output": ctx, "image_encoder_output": None}
        outs = []
        for parallel in (False, True):
            model = WanModel.from_weight_dict(cfg, W)
            sched = WanScheduler(cfg, device=dev)
            sched.prepare()
            model.set_scheduler(sched)
            if parallel:
                mode = U.parallelize_wan_cfg(model, 3 * 8 * 8, lib.fmha, sp="nccl")
                assert mode == "cfg2", mode
            sched.step_pre(0)
            model.infer(inputs)
            outs.append(sched.noise_pred.clone())
        torch.cuda.synchronize()
        if rank == 0:
            torch.save({"equal": bool(torch.equal(outs[0], outs[1])), "max": float((outs[0] - outs[1]).abs().max())}, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_cfg_parallel_equals_serial_cfg(tmp_path):
    """cond on rank 0, uncond on rank 1, predictions exchanged: bit-identical to the serial two-pass CFG on one GPU."""
    import torch.multiprocessing as mp

    out = str(tmp_path / "res_c.pt")
    mp.spawn(_run_cfg_rank, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    print("cfg-parallel vs serial:", r)
    assert r["equal"], r
