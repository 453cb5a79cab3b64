"""Flow-matching UniPC (order 2, B(h) = expm1(h)) sampler with the reference scheduler's interface
(`prepare`, `step_pre(step_index)`, `step_post()`, fields `latents`, `timesteps`, `sigmas`, `noise_pred`;
lightx2v/models/schedulers/scheduler.py:5-21, lightx2v/models/schedulers/wan/scheduler.py:9-360) and its arithmetic:
sigma schedule linspace(sigma_max, sigma_min, steps+1)[:-1] with shift s*sig/(1+(s-1)*sig) (:65-94), x0 = x - sigma*v (:96-116),
UniPC predictor / corrector in fp32 with fp32 CPU scalars (:130-320), corrector from the second step on (:322-360).
The step-distill variant (lightx2v/models/schedulers/wan/step_distill/scheduler.py:8-56) predicts x0 and re-noises it
to the next sigma of a fixed timestep list.  O(latent) elementwise work: stays in torch, on whatever device the latents live on."""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch


class WanScheduler:
    def __init__(self, config, device: Optional[torch.device] = None):
        self.config = config
        self.device = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
        self.infer_steps = config["infer_steps"]
        self.sample_shift = config.get("sample_shift", 5.0)
        self.num_train_timesteps = 1000
        self.solver_order = 2
        self.disable_corrector: List[int] = []
        self.step_index = 0
        self.latents = None
        self.noise_pred = None
        self.flag_df = False
        self.bf16_step_pre = config.get("dtype_mode", "BF16") == "BF16"

    # ------------------------------------------------------------------ setup
    def prepare(self, image_encoder_output=None):
        self.generator = torch.Generator(device=self.device)
        self.generator.manual_seed(self.config["seed"])
        ts = self.config["target_shape"]
        self.latents = torch.randn(ts[0], ts[1], ts[2], ts[3], dtype=torch.float32, device=self.device, generator=self.generator)
        ps = self.config.get("patch_size", (1, 2, 2))
        if self.config.get("task", "t2v") == "i2v" and "lat_h" in self.config:      # scheduler.py:32-33 (floor, from the VAE grid)
            frames = (self.config["target_video_length"] - 1) // self.config.get("vae_stride", (4, 8, 8))[0] + 1
            self.seq_len = frames * self.config["lat_h"] * self.config["lat_w"] // (ps[1] * ps[2])
        else:                                                                       # scheduler.py:30-31
            self.seq_len = math.ceil((ts[2] * ts[3]) / (ps[1] * ps[2]) * ts[1])
        self._seed_default_generator()
        alphas = np.linspace(1, 1 / self.num_train_timesteps, self.num_train_timesteps)[::-1].copy()
        sig = torch.from_numpy(1.0 - alphas).to(dtype=torch.float32)
        self.sigma_min = sig[-1].item()
        self.sigma_max = sig[0].item()
        self.set_timesteps(self.infer_steps, shift=self.sample_shift)

    def _seed_default_generator(self):
        """Sequence / CFG parallel runs replicate the latents on every rank; any draw from the process DEFAULT generator (the step-distill
        re-noising, step_distill/scheduler.py:53) must therefore be identical across ranks.  The reference gets this from its runner's
        seed_all(config.seed) at start-up (lightx2v/utils/utils.py seed_all); a stand-alone user of this scheduler gets it here."""
        try:
            import torch.distributed as dist
            multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        except Exception:
            multi = False
        if multi:
            torch.manual_seed(self.config["seed"])
            if self.device.type == "cuda":
                torch.cuda.manual_seed(self.config["seed"])

    def set_timesteps(self, infer_steps, shift=1.0):
        sig = np.linspace(self.sigma_max, self.sigma_min, infer_steps + 1).copy()[:-1]
        sig = shift * sig / (1 + (shift - 1) * sig)
        timesteps = sig * self.num_train_timesteps
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0]]).astype(np.float32))          # CPU fp32 scalars
        self.timesteps = torch.from_numpy(timesteps).to(device=self.device, dtype=torch.int64)
        self.model_outputs = [None] * self.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self.this_order = None

    # ------------------------------------------------------------------ per step
    def step_pre(self, step_index):
        self.step_index = step_index
        if self.bf16_step_pre:
            self.latents = self.latents.to(dtype=torch.bfloat16)

    @staticmethod
    def _lambda(sigma):
        return torch.log(1 - sigma) - torch.log(sigma)

    def _coeffs(self, i_t, i_s0, order, hist_idx):
        """Shared algebra of predictor and corrector: returns (ratio, alpha_t, h_phi_1, B_h, rks, b) as fp32 CPU scalars."""
        sigma_t, sigma_s0 = self.sigmas[i_t], self.sigmas[i_s0]
        alpha_t = 1 - sigma_t
        h = self._lambda(sigma_t) - self._lambda(sigma_s0)
        rks = []
        for si in hist_idx[: order - 1]:
            rks.append((self._lambda(self.sigmas[si]) - self._lambda(sigma_s0)) / h)
        rks.append(1.0)
        hh = -h
        h_phi_1 = torch.expm1(hh)
        B_h = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        b, fact = [], 1
        for i in range(1, order + 1):
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        return sigma_t / sigma_s0, alpha_t, h_phi_1, B_h, rks, b

    def _predict(self, sample, order):
        m0 = self.model_outputs[-1]
        i = self.step_index
        ratio, alpha_t, h_phi_1, B_h, rks, _ = self._coeffs(i + 1, i, order, [i - 1])
        dev = sample.device
        x_t = ratio.to(dev) * sample - (alpha_t * h_phi_1).to(dev) * m0
        if order == 2:
            D1 = (self.model_outputs[-2] - m0) / rks[0].to(dev)
            x_t = x_t - (alpha_t * B_h).to(dev) * (0.5 * D1)
        return x_t.to(sample.dtype)

    def _correct(self, model_t, last_sample, order):
        m0 = self.model_outputs[-1]
        i = self.step_index
        ratio, alpha_t, h_phi_1, B_h, rks, b = self._coeffs(i, i - 1, order, [i - 2])
        dev = last_sample.device
        x_t = ratio.to(dev) * last_sample - (alpha_t * h_phi_1).to(dev) * m0
        if order == 1:
            corr = 0.5 * (model_t - m0)
        else:
            rk = torch.stack([torch.as_tensor(r, dtype=torch.float32) for r in rks])
            R = torch.stack([torch.pow(rk, k) for k in range(order)])
            rhos = torch.linalg.solve(R, torch.stack([torch.as_tensor(v, dtype=torch.float32) for v in b])).to(torch.float32)
            D1 = (self.model_outputs[-2] - m0) / rks[0].to(dev)
            corr = rhos[0].to(dev) * D1 + rhos[1].to(dev) * (model_t - m0)
        x_t = x_t - (alpha_t * B_h).to(dev) * corr
        return x_t.to(last_sample.dtype)

    def step_post(self):
        v = self.noise_pred.to(torch.float32)
        sample = self.latents.to(torch.float32)
        x0 = sample - self.sigmas[self.step_index].to(sample.device) * v          # convert_model_output
        if self.step_index > 0 and (self.step_index - 1) not in self.disable_corrector and self.last_sample is not None:
            sample = self._correct(x0, self.last_sample, self.this_order)
        self.model_outputs = self.model_outputs[1:] + [x0]
        order = min(self.solver_order, len(self.timesteps) - self.step_index)
        self.this_order = min(order, self.lower_order_nums + 1)
        self.last_sample = sample
        self.latents = self._predict(sample, self.this_order)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1

    def clear(self):
        pass


class WanStepDistillScheduler(WanScheduler):
    """4-step distilled sampler (config 3) — lightx2v/models/schedulers/wan/step_distill/scheduler.py:8-56:
    sigma grid linspace(1, 0, 1001)[:-1] shifted, picked at `denoising_step_list`; each step predicts x0 = x - sigma*v and,
    except after the last step, re-noises it to the next sigma: x = (1 - s') * x0 + s' * randn."""

    def __init__(self, config, device=None):
        super().__init__(config, device)
        self.denoising_step_list = config.get("denoising_step_list", [1000, 750, 500, 250])
        self.infer_steps = len(self.denoising_step_list)
        self.sigma_max, self.sigma_min = 1.0, 0.0

    def prepare(self, image_encoder_output=None):
        super().prepare(image_encoder_output)
        self.sigma_max, self.sigma_min = 1.0, 0.0
        sig = torch.linspace(1.0, 0.0, self.num_train_timesteps + 1)[:-1]
        sig = self.sample_shift * sig / (1 + (self.sample_shift - 1) * sig)
        idx = [self.num_train_timesteps - t for t in self.denoising_step_list]
        self.timesteps = (sig * self.num_train_timesteps)[idx].to(self.device)
        self.sigmas = sig[idx].to("cpu")

    def step_post(self):
        v = self.noise_pred.to(torch.float32)
        sigma = self.sigmas[self.step_index].item()
        x0 = self.latents.to(torch.float32) - sigma * v
        if self.step_index < self.infer_steps - 1:
            s1 = self.sigmas[self.step_index + 1].item()
            noise = torch.randn_like(x0)                 # the reference draws from the device's DEFAULT generator here (:53), not the seeded one
            x0 = ((1 - s1) * x0 + s1 * noise).type_as(noise)
        self.latents = x0.to(self.latents.dtype)
