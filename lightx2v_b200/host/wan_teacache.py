"""TeaCache around the B200 block stack: skip the 40 blocks of a forward when the timestep-embedding input has moved little since the
last computed step, and re-apply the cached residual instead.

Mirrors WanTransformerInferTeaCaching (lightx2v/models/networks/wan/infer/feature_caching/transformer_infer.py:9-171) decision for
decision: separate state for the conditional ("even") and unconditional ("odd") pass, warm-up `ret_steps` / `cutoff_steps`, relative L1
change of the modulation input rescaled by the config's polynomial (np.poly1d(coefficients)), accumulated until it crosses
`teacache_thresh`.  The decision needs one scalar on the host per forward (it gates which kernels are launched), so one .item() per
forward is inherent to the algorithm; everything else stays on the device.  Config keys as the reference's configs/caching/teacache/*.json:
`teacache_thresh`, `coefficients` (two rows), `use_ret_steps`, `infer_steps`, `enable_cfg`."""
from __future__ import annotations

import numpy as np
import torch

from .wan_infer import WanTransformerInfer


class WanTransformerInferTeaCaching(WanTransformerInfer):
    def __init__(self, config):
        super().__init__(config)
        self.cnt = 0
        self.teacache_thresh = config["teacache_thresh"]
        self.accumulated_rel_l1_distance_even = 0
        self.previous_e0_even = None
        self.previous_residual_even = None
        self.accumulated_rel_l1_distance_odd = 0
        self.previous_e0_odd = None
        self.previous_residual_odd = None
        self.use_ret_steps = config["use_ret_steps"]
        if self.use_ret_steps:                                              # :20-27
            self.coefficients = config["coefficients"][0]
            self.ret_steps = 5 * 2
            self.cutoff_steps = config["infer_steps"] * 2
        else:
            self.coefficients = config["coefficients"][1]
            self.ret_steps = 1 * 2
            self.cutoff_steps = config["infer_steps"] * 2 - 2
        self.scheduler = None
        self._rescale = np.poly1d(self.coefficients)

    def set_scheduler(self, scheduler):
        self.scheduler = scheduler
        n = scheduler.infer_steps
        if not hasattr(scheduler, "caching_records"):
            scheduler.caching_records = [True] * n                          # BaseScheduler.__init__, schedulers/scheduler.py:11
        if not hasattr(scheduler, "caching_records_2"):
            scheduler.caching_records_2 = [True] * n

    # ------------------------------------------------------------------ decision (:30-82)
    def calculate_should_calc(self, embed, embed0) -> bool:
        modulated_inp = embed0 if self.use_ret_steps else embed
        tag = "even" if self.infer_conditional else "odd"
        prev = getattr(self, "previous_e0_" + tag)
        acc = getattr(self, "accumulated_rel_l1_distance_" + tag)
        if self.cnt < self.ret_steps or self.cnt >= self.cutoff_steps:
            should_calc, acc = True, 0
        else:
            rel = ((modulated_inp - prev).abs().mean() / prev.abs().mean()).cpu().item()
            acc += self._rescale(rel)
            if acc < self.teacache_thresh:
                should_calc = False
            else:
                should_calc, acc = True, 0
        setattr(self, "accumulated_rel_l1_distance_" + tag, acc)
        setattr(self, "previous_e0_" + tag, modulated_inp.clone())
        return should_calc

    # ------------------------------------------------------------------ forward (:84-118)
    def infer(self, weights, grid_sizes, embed, x, embed0, seq_lens, freqs, context, audio_dit_blocks=None):
        index = self.scheduler.step_index
        records = self.scheduler.caching_records if self.infer_conditional else self.scheduler.caching_records_2
        if index <= self.scheduler.infer_steps - 1:
            records[index] = self.calculate_should_calc(embed, embed0)
        if records[index]:
            x = self.infer_calculating(weights, grid_sizes, embed, x, embed0, seq_lens, freqs, context)
        else:
            x = self.infer_using_cache(x)
        if self.config.get("enable_cfg", False):
            self.switch_status()
        self.cnt += 1
        return x

    def infer_calculating(self, weights, grid_sizes, embed, x, embed0, seq_lens, freqs, context):
        ori_x = x.clone()                                                   # the block stack updates x in place
        x = super().infer(weights, grid_sizes, embed, x, embed0, seq_lens, freqs, context)
        residual = x - ori_x
        if self.infer_conditional:
            self.previous_residual_even = residual
        else:
            self.previous_residual_odd = residual
        return x

    def infer_using_cache(self, x):
        x.add_(self.previous_residual_even if self.infer_conditional else self.previous_residual_odd)
        return x

    def clear(self):
        self.previous_residual_even = self.previous_residual_odd = None
        self.previous_e0_even = self.previous_e0_odd = None
