"""One CUDA graph per denoise step: `scheduler.step_pre -> WanModel.infer (cond + uncond forwards, CFG combine) ->
scheduler.step_post` captured once per step KIND and replayed, with a device-resident UniPC scheduler.

The reference's loop body (lightx2v/models/runners/default_runner.py:97-114) launches every kernel from Python and its scheduler computes
the UniPC coefficients as fp32 CPU scalars that are copied to the device inside every step (lightx2v/models/schedulers/wan/scheduler.py:
130-360).  Here
  * `WanSchedulerDevice` computes the same fp32 scalars ONCE per schedule on the host (same code path as WanScheduler._coeffs) into a
    [steps, 12] device table; before a step one row is copied into a static 12-float buffer (a 48-byte device-to-device copy), and the
    update reads its coefficients from there.  Latents, last sample and the two model-output history slots live in static buffers that
    are updated in place, so a captured graph is valid for every step of its kind;
  * a step's kind is (corrector order 0 / 1 / 2, predictor order 1 / 2): at most four graphs per schedule (first step, second step, steady
    state, last step);
  * the sinusoidal timestep embedding of the current step is one row of a table built once (the reference recomputes it in fp64 every
    step, wan/infer/utils.py:161-172); the row is copied into a static buffer next to the coefficients.
Arithmetic is expression-for-expression the eager scheduler's (tests/test_gpu_model.py requires bit-identical latents after several steps).
Single-GPU, bf16 / fp8 / nvfp4 linears.  Sequence-parallel runs keep the eager loop (their cross-GPU barriers are stream-ordered host calls).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from .wan_model import sinusoidal_embedding_1d
from .wan_scheduler import WanScheduler

# columns of the coefficient table
SIGMA, P_RATIO, P_AH, P_AB, P_RK0, C_RATIO, C_AH, C_AB, C_RK0, C_RHO0, C_RHO1 = range(11)
NCOEF = 12


class WanSchedulerDevice(WanScheduler):
    def prepare(self, image_encoder_output=None):
        super().prepare(image_encoder_output)
        self._build_tables()
        lat = self.latents
        self.s_lat = lat.clone()                                  # fp32 latents (state)
        self.s_last = torch.zeros_like(lat)
        self.s_m0 = torch.zeros_like(lat)                         # newest converted model output (x0)
        self.s_m1 = torch.zeros_like(lat)                         # the one before
        self.c = torch.zeros(NCOEF, dtype=torch.float32, device=self.device)
        self.t_embed_static = torch.zeros(1, self.config.get("freq_dim", 256), dtype=torch.bfloat16, device=self.device)

    def set_timesteps(self, infer_steps, shift=1.0):
        super().set_timesteps(infer_steps, shift)
        if hasattr(self, "c"):
            self._build_tables()

    # ------------------------------------------------------------------ host side, once per schedule
    def _build_tables(self):
        n = len(self.timesteps)
        rows = torch.zeros(n, NCOEF, dtype=torch.float32)
        kinds = []
        lower, this_order = 0, None
        for i in range(n):
            self.step_index = i
            corr = 0
            if i > 0 and (i - 1) not in self.disable_corrector:
                corr = this_order
                ratio, alpha_t, h_phi_1, B_h, rks, b = self._coeffs(i, i - 1, corr, [i - 2])
                rows[i, C_RATIO], rows[i, C_AH], rows[i, C_AB] = ratio, alpha_t * h_phi_1, alpha_t * B_h
                if corr == 2:
                    rk = torch.stack([torch.as_tensor(r, dtype=torch.float32) for r in rks])
                    R = torch.stack([torch.pow(rk, k) for k in range(corr)])
                    rhos = torch.linalg.solve(R, torch.stack([torch.as_tensor(v, dtype=torch.float32) for v in b])).to(torch.float32)
                    rows[i, C_RK0], rows[i, C_RHO0], rows[i, C_RHO1] = rks[0], rhos[0], rhos[1]
            order = min(self.solver_order, n - i)
            this_order = min(order, lower + 1)
            ratio, alpha_t, h_phi_1, B_h, rks, _ = self._coeffs(i + 1, i, this_order, [i - 1])
            rows[i, SIGMA], rows[i, P_RATIO], rows[i, P_AH], rows[i, P_AB] = self.sigmas[i], ratio, alpha_t * h_phi_1, alpha_t * B_h
            if this_order == 2:
                rows[i, P_RK0] = rks[0]
            kinds.append((corr, this_order))
            if lower < self.solver_order:
                lower += 1
        self.step_index = 0
        self.table = rows.to(self.device)
        self.kinds = kinds
        self.t_table = sinusoidal_embedding_1d(self.config.get("freq_dim", 256), self.timesteps.flatten().cpu()).to(self.device)

    def kind(self, i: int) -> Tuple[int, int]:
        return self.kinds[i]

    def load_step(self, i: int):
        """Outside the graph: select step i (two tiny device-to-device copies on the current stream)."""
        self.step_index = i
        self.c.copy_(self.table[i])
        self.t_embed_static.copy_(self.t_table[i:i + 1])

    # ------------------------------------------------------------------ device side (capturable)
    def step_pre(self, step_index=None):
        if step_index is not None:
            self.step_index = step_index
        self.latents = self.s_lat.to(dtype=torch.bfloat16) if self.bf16_step_pre else self.s_lat

    def step_post_device(self, kind: Tuple[int, int]):
        """WanScheduler.step_post with every scalar read from `self.c` and every piece of state updated in place."""
        corr, order = kind
        c = self.c
        v = self.noise_pred.to(torch.float32)
        sample = self.latents.to(torch.float32)
        x0 = sample - c[SIGMA] * v
        if corr:
            m0 = self.s_m0
            x_t = c[C_RATIO] * self.s_last - c[C_AH] * m0
            if corr == 1:
                cr = 0.5 * (x0 - m0)
            else:
                D1 = (self.s_m1 - m0) / c[C_RK0]
                cr = c[C_RHO0] * D1 + c[C_RHO1] * (x0 - m0)
            sample = (x_t - c[C_AB] * cr).to(self.s_last.dtype)
        self.s_m1.copy_(self.s_m0)
        self.s_m0.copy_(x0)
        self.s_last.copy_(sample)
        m0 = self.s_m0
        x_t = c[P_RATIO] * sample - c[P_AH] * m0
        if order == 2:
            D1 = (self.s_m1 - m0) / c[P_RK0]
            x_t = x_t - c[P_AB] * (0.5 * D1)
        self.s_lat.copy_(x_t.to(sample.dtype))

    def step_post(self):
        self.step_post_device(self.kinds[self.step_index])
        self.latents = self.s_lat


class GraphedDenoiser:
    """`step(i)` = one denoise step of `model` (a host.wan_model.WanModel with a WanSchedulerDevice) as a CUDA-graph replay."""

    def __init__(self, model, scheduler: WanSchedulerDevice, inputs: Dict):
        if model.pre_process is not None or model.cfg_parallel is not None:
            raise RuntimeError("GraphedDenoiser: the sequence / CFG parallel paths keep the eager loop")
        self.model, self.sched, self.inputs = model, scheduler, inputs
        self.graphs: Dict[Tuple[int, int], torch.cuda.CUDAGraph] = {}
        self.kernel_nodes: Dict[Tuple[int, int], int] = {}      # library kernel launches recorded into each graph
        self.replayed_launches = 0                             # library kernels executed through replays so far
        self.pool = None
        model.pre_infer.use_static_t_embed = True

    def _body(self, kind):
        self.sched.step_pre()
        self.model.infer(self.inputs)
        self.sched.step_post_device(kind)

    def _state(self):
        s = self.sched
        return [t.clone() for t in (s.s_lat, s.s_last, s.s_m0, s.s_m1)]

    def _restore(self, saved):
        s = self.sched
        for dst, src in zip((s.s_lat, s.s_last, s.s_m0, s.s_m1), saved):
            dst.copy_(src)

    def step(self, i: int):
        s = self.sched
        kind = s.kind(i)
        s.load_step(i)
        g = self.graphs.get(kind)
        if g is None:
            # first step of this kind: run it once eagerly on a side stream (fills the rope / text-K,V / workspace caches, which may
            # synchronise and therefore cannot happen under capture), put the state back, then capture
            saved = self._state()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._body(kind)
            torch.cuda.current_stream().wait_stream(side)
            self._restore(saved)
            from .. import lib
            g = torch.cuda.CUDAGraph()
            n0 = lib.launch_count()
            with torch.cuda.graph(g, pool=self.pool):
                self._body(kind)
            self.kernel_nodes[kind] = lib.launch_count() - n0
            if self.pool is None:
                self.pool = g.pool()
            self.graphs[kind] = g
        g.replay()
        self.replayed_launches += self.kernel_nodes[kind]
        s.latents = s.s_lat
