"""world_size = 2 `gloo` tests (CPU) of the Ulysses host logic: shard / pad / gather, packed all-to-all layouts, head
scatter + sequence gather around an injected attention function (the oracle's SDPA — the CUDA FMHA is exercised by the -m gpu tests)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import wan_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_attn(q, k, v, out=None):
    o = O.attn_apply(q.float(), k.float(), v.float()).view(q.shape[0], q.shape[1], q.shape[2]).to(q.dtype)
    if out is not None:
        out.copy_(o)
        return out
    return o


def _worker(rank, world, port, S, H, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lightx2v_b200.host import ulysses as U

        torch.manual_seed(0)
        d = 128
        x = torch.randn(S, 3 * H * d).to(torch.bfloat16)          # a fused QKV activation for all S tokens
        # ---- shard / gather round trip with padding
        xs = U.pre_process(x, rank, world)
        s = U.shard_rows(S, world)
        assert xs.shape == (s, 3 * H * d)
        assert torch.equal(U.post_process(xs, S), x)
        # ---- sequence-parallel attention == attention over the full sequence
        q, k, v = (xs[:, i * H * d:(i + 1) * H * d].unflatten(1, (H, d)) for i in range(3))     # strided views, like the QKV buffer
        att = U.UlyssesAttention(_oracle_attn, total_rows=S)
        o_local = att(q=q, k=k, v=v)
        o_full = U.post_process(o_local.reshape(s, H * d), S)
        fq, fk, fv = (x[:, i * H * d:(i + 1) * H * d].unflatten(1, (H, d)) for i in range(3))
        ref = _oracle_attn(fq, fk, fv).reshape(S, H * d)
        results[rank] = (torch.equal(o_full, ref), float((o_full.float() - ref.float()).abs().max()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("S,H", [(64, 4), (45, 2)])       # 45 tokens: pad to 46, the pad row must not act as a key
def test_ulysses_matches_full_attention_world2(S, H):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, S, H, results), nprocs=world, join=True)
    for r in range(world):
        same, err = results[r]
        assert same, f"rank {r}: max err {err}"


def test_pack_unpack_layouts():
    from lightx2v_b200.host import ulysses as U

    s, H, d, P = 6, 4, 128, 2
    q, k, v = (torch.arange(s * H * d, dtype=torch.float32).view(s, H, d) + o for o in (0, 1e6, 2e6))
    send = U.pack_qkv_for_a2a(q, k, v, P)
    assert send.shape == (P, s, 3, H // P, d)
    for p in range(P):
        for i, t in enumerate((q, k, v)):
            assert torch.equal(send[p, :, i], t[:, p * (H // P):(p + 1) * (H // P)])
    recv = torch.stack([q[:, p * (H // P):(p + 1) * (H // P)] for p in range(P)])      # [P, s, H/P, d]
    assert torch.equal(U.unpack_out_from_a2a(recv), q)


def test_heads_must_divide_world():
    from lightx2v_b200.host import ulysses as U

    class FakeAtt(U.UlyssesAttention):
        def __init__(self):
            self.world, self.rank, self.group, self.total_rows, self._bufs, self.attn = 3, 0, None, 8, {}, None

    with pytest.raises(ValueError):
        FakeAtt()(q=torch.zeros(4, 4, 128), k=torch.zeros(4, 4, 128), v=torch.zeros(4, 4, 128))


def _hunyuan_worker(rank, world, port, Li, Lt, valid, H, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lightx2v_b200.host import ulysses as U

        torch.manual_seed(1)
        d = 128
        img = torch.randn(Li, 3, H, d).to(torch.bfloat16)           # post-norm/rope q, k, v of all image tokens
        txt = torch.randn(Lt, 3, H, d).to(torch.bfloat16)           # text tokens: replicated on every rank
        s = Li // world
        local = torch.cat((img[rank * s:(rank + 1) * s], txt), 0)
        bounds = [0, Li + valid, Li + Lt]
        att = U.HunyuanUlyssesAttention(_oracle_attn)
        out = torch.empty(s + Lt, H, d, dtype=torch.bfloat16)
        att(local, s, bounds, out)
        # reference: joint attention over [all image ; text] with the two varlen segments (hunyuan_oracle.varlen_attention)
        full = torch.cat((img, txt), 0)
        ref = torch.empty(Li + Lt, H, d, dtype=torch.bfloat16)
        for a, b in zip(bounds[:-1], bounds[1:]):
            ref[a:b] = _oracle_attn(full[a:b, 0], full[a:b, 1], full[a:b, 2])
        ok_img = torch.equal(out[:s], ref[rank * s:(rank + 1) * s])
        ok_txt = torch.equal(out[s:], ref[Li:])
        results[rank] = (ok_img and ok_txt, float((out[:s].float() - ref[rank * s:(rank + 1) * s].float()).abs().max()))
    finally:
        dist.destroy_process_group()


def test_hunyuan_ulysses_matches_joint_attention_world2():
    """Image tokens sharded, text replicated, heads scattered: equals the joint [image ; text] attention with the padded-text segment."""
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        results = m.dict()
        mp.spawn(_hunyuan_worker, args=(world, port, 48, 8, 5, 4, results), nprocs=world, join=True)
        for r in range(world):
            ok, err = results[r]
            assert ok, (r, err)


def _cfg_worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lightx2v_b200.host import ulysses as U
        from lightx2v_b200.host.wan_model import WanModel

        class Sched:
            noise_pred = None

        model = object.__new__(WanModel)
        model.config = dict(enable_cfg=True, sample_guide_scale=5.0)
        model.scheduler, model.cfg_parallel, model.pre_process, model.post_process = Sched(), None, None, None
        calls = []

        def forward(inputs, is_cond):                    # stands for the whole block stack: a function of the branch only
            calls.append(is_cond)
            g = torch.Generator().manual_seed(1 if is_cond else 2)
            return torch.randn(16, 3, 8, 8, generator=g)

        model._forward = forward
        WanModel.infer(model, {})                        # serial CFG: cond then uncond on this rank
        serial = model.scheduler.noise_pred.clone()
        assert calls == [True, False]
        del calls[:]
        mode = U.parallelize_wan_cfg(model, 3 * 8 * 8, attention_fn=_oracle_attn, sp="nccl")
        WanModel.infer(model, {})
        results[rank] = (mode, list(calls), torch.equal(model.scheduler.noise_pred, serial))
    finally:
        dist.destroy_process_group()


def test_cfg_parallel_world2_runs_one_branch_per_rank_and_equals_serial_cfg():
    """parallelize_wan_cfg on two ranks: rank 0 runs only the conditional forward, rank 1 only the unconditional one, the predictions are
    exchanged and both ranks hold `uncond + g (cond - uncond)` bit-identical to the serial evaluation (wan/model.py:203-218)."""
    world = 2
    with mp.Manager() as m:
        results = m.dict()
        mp.spawn(_cfg_worker, args=(world, _free_port(), results), nprocs=world, join=True)
        assert results[0] == ("cfg2", [True], True) and results[1] == ("cfg2", [False], True), dict(results)
