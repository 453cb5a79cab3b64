"""Wan DiT per-block op tree with the reference's structure and checkpoint key names
(lightx2v/models/networks/wan/weights/transformer_weights.py:14-366):
  WanTransformerWeights.blocks[i].compute_phases = [WanModulation, WanSelfAttention, WanCrossAttention, WanFFN].
Op classes are looked up in this package's registries by the same config strings the reference uses
(`mm_config.mm_type`, `self_attn_1_type`, `cross_attn_1_type`, `cross_attn_2_type`)."""
from __future__ import annotations

from . import ops  # noqa: F401  (registers the op classes)
from .registry import ATTN_KEY, ATTN_WEIGHT_REGISTER, LN_WEIGHT_REGISTER, MM_KEY, MM_WEIGHT_REGISTER, RMS_WEIGHT_REGISTER, TENSOR_REGISTER
from .weight_module import WeightModule, WeightModuleList


class WanTransformerWeights(WeightModule):
    def __init__(self, config):
        super().__init__()
        self.blocks_num = config["num_layers"]
        self.task = config["task"]
        self.config = config
        self.mm_type = (config.get("mm_config") or {}).get("mm_type", MM_KEY)
        self.blocks = WeightModuleList([WanTransformerAttentionBlock(i, self.task, self.mm_type, self.config) for i in range(self.blocks_num)])
        self.add_module("blocks", self.blocks)

    def clear(self):
        for block in self.blocks:
            for phase in block.compute_phases:
                phase.clear()


class WanTransformerAttentionBlock(WeightModule):
    def __init__(self, block_index, task, mm_type, config):
        super().__init__()
        self.block_index = block_index
        self.mm_type = mm_type
        self.task = task
        self.config = config
        self.compute_phases = WeightModuleList(
            [
                WanModulation(block_index, task, mm_type, config),
                WanSelfAttention(block_index, task, mm_type, config),
                WanCrossAttention(block_index, task, mm_type, config),
                WanFFN(block_index, task, mm_type, config),
            ]
        )
        self.add_module("compute_phases", self.compute_phases)


class _Phase(WeightModule):
    def __init__(self, block_index, task, mm_type, config):
        super().__init__()
        self.block_index = block_index
        self.mm_type = mm_type
        self.task = task
        self.config = config

    def _mm(self, name, key):
        p = f"blocks.{self.block_index}.{key}"
        self.add_module(name, MM_WEIGHT_REGISTER[self.mm_type](p + ".weight", p + ".bias"))

    def _rms(self, name, key):
        self.add_module(name, RMS_WEIGHT_REGISTER["sgl-kernel"](f"blocks.{self.block_index}.{key}.weight"))


class WanModulation(_Phase):
    """transformer_weights.py:90-113"""

    def __init__(self, block_index, task, mm_type, config):
        super().__init__(block_index, task, mm_type, config)
        self.add_module("modulation", TENSOR_REGISTER["Default"](f"blocks.{block_index}.modulation"))


class WanSelfAttention(_Phase):
    """transformer_weights.py:116-209"""

    def __init__(self, block_index, task, mm_type, config):
        super().__init__(block_index, task, mm_type, config)
        self.add_module("norm1", LN_WEIGHT_REGISTER["Default"]())
        for nm in ("q", "k", "v", "o"):
            self._mm(f"self_attn_{nm}", f"self_attn.{nm}")
        self._rms("self_attn_norm_q", "self_attn.norm_q")
        self._rms("self_attn_norm_k", "self_attn.norm_k")
        self.add_module("self_attn_1", ATTN_WEIGHT_REGISTER[config.get("self_attn_1_type", ATTN_KEY)]())


class WanCrossAttention(_Phase):
    """transformer_weights.py:212-312"""

    def __init__(self, block_index, task, mm_type, config):
        super().__init__(block_index, task, mm_type, config)
        p = f"blocks.{block_index}"
        self.add_module("norm3", LN_WEIGHT_REGISTER["Default"](f"{p}.norm3.weight", f"{p}.norm3.bias"))
        for nm in ("q", "k", "v", "o"):
            self._mm(f"cross_attn_{nm}", f"cross_attn.{nm}")
        self._rms("cross_attn_norm_q", "cross_attn.norm_q")
        self._rms("cross_attn_norm_k", "cross_attn.norm_k")
        self.add_module("cross_attn_1", ATTN_WEIGHT_REGISTER[config.get("cross_attn_1_type", ATTN_KEY)]())
        if task == "i2v":
            self._mm("cross_attn_k_img", "cross_attn.k_img")
            self._mm("cross_attn_v_img", "cross_attn.v_img")
            self._rms("cross_attn_norm_k_img", "cross_attn.norm_k_img")
            self.add_module("cross_attn_2", ATTN_WEIGHT_REGISTER[config.get("cross_attn_2_type", ATTN_KEY)]())


class WanFFN(_Phase):
    """transformer_weights.py:315-366"""

    def __init__(self, block_index, task, mm_type, config):
        super().__init__(block_index, task, mm_type, config)
        self.add_module("norm2", LN_WEIGHT_REGISTER["Default"]())
        self._mm("ffn_0", "ffn.0")
        self._mm("ffn_2", "ffn.2")
