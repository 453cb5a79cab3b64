"""CPU tests of the HunyuanVideo VAE host logic (host/hunyuan_vae.py) around an injected per-tile decoder: the (temporal, row, column)
tile traversal and the linear blends against the oracle's restatement of AutoencoderKLCausal3D's tiling (oracle/hunyuan_vae_oracle.py,
itself pinned to the real class by tests/test_oracle_golden.py), and the tile-parallel decode_dist on a world_size-2 `gloo` group against
the single-process decode (plan pass -> own tiles -> exchange -> blends).  The CUDA per-tile decoder is exercised by the -m gpu tests."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import hunyuan_vae_oracle as HV

CFG = {"block_out_channels": (128, 256, 512, 512), "sample_size": 64, "sample_tsize": 16, "tile_overlap_factor": 0.25, "scaling_factor": 1.0}
# -> 8 x 8 latent tiles (step 6, blend over 16 pixels), temporal tiles of 4 + 1 latent frames (step 3, blend over 4 frames)


def _stub_tile(z: torch.Tensor) -> torch.Tensor:
    """[16, t, h, w] -> [3, 1 + 4 (t - 1), 8 h, 8 w] fp32: causal-VAE shaped, and a function of the WHOLE tile (its mean enters every
    pixel, as GroupNorm statistics do), so a tile decoded from the wrong window or blended in the wrong order changes the result."""
    x = z[:3] * 0.5 + z[3:6].mean() + 0.1 * z.mean()
    x = x.repeat_interleave(8, -1).repeat_interleave(8, -2)
    return torch.cat((x[:, :1], x[:, 1:].repeat_interleave(4, 1)), 1).float().contiguous()


class _StubDecoder:
    cfg = CFG
    device = torch.device("cpu")

    def decode_tile(self, z):
        return _stub_tile(z)


def _vae():
    from lightx2v_b200.host.hunyuan_vae import HunyuanVAEB200

    v = object.__new__(HunyuanVAEB200)
    v.decoder, v.cfg, v.device = _StubDecoder(), CFG, torch.device("cpu")
    v.tile_sample, v.tile_latent = CFG["sample_size"], CFG["sample_size"] // 8
    v.tile_tsample, v.tile_tlatent = CFG["sample_tsize"], CFG["sample_tsize"] // 4
    v.overlap = CFG["tile_overlap_factor"]
    v._dist, v._tile_counter = None, 0
    return v


def _latents(T, H, W):
    return torch.randn(1, 16, T, H, W, generator=torch.Generator().manual_seed(11))


def test_tile_traversal_and_blends_match_the_oracle(monkeypatch):
    monkeypatch.setattr(HV, "tile_decode", lambda W, z, cfg: _stub_tile(z[0]).unsqueeze(0))
    v = _vae()
    for shape in ((9, 12, 20), (3, 12, 20), (4, 8, 8), (9, 8, 8), (7, 9, 7)):      # temporal + spatial, spatial only, one tile, temporal only, ragged
        lat = _latents(*shape)
        got = v.decode_device(lat)
        ref = HV.decode(None, lat.clone(), CFG)
        assert got.shape == ref.shape == (1, 3, 1 + 4 * (shape[0] - 1), 8 * shape[1], 8 * shape[2])
        assert torch.equal(got, ref), (shape, (got - ref).abs().max())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        v = _vae()
        calls = []
        inner = v.decoder.decode_tile
        v.decoder.decode_tile = lambda z: (calls.append(1), inner(z))[1]
        lat = _latents(9, 12, 20)                       # 3 temporal x 2 x 4 spatial tiles = 24 tiles
        got = v.decode_dist(lat, to_cpu=True)
        mine = len(calls)
        ref = v.decode_device(lat)                      # every tile locally
        total = len(calls) - mine
        results[rank] = (torch.equal(got, ref), mine, total, v._dist is None and not v._plan and not v._done)
    finally:
        dist.destroy_process_group()


def test_tile_parallel_decode_world2_equals_the_single_process_decode():
    world = 2
    with mp.Manager() as m:
        results = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
        for r in range(world):
            same, mine, total, clean = results[r]
            assert same and clean, (r, results[r])
            assert total == 24 and mine == 12          # tile k belongs to rank k mod P: each rank decoded exactly half, once
