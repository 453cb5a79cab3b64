"""String-keyed op registries with the reference's semantics (lightx2v/utils/registry_factory.py:1-56):
`@REGISTER("key")` decorates a class, duplicate keys raise, `REGISTER[key]` looks a class up, `REGISTER[key] = cls`
overrides without a check (that asymmetry is what lets a plugin replace the hard-coded "sgl-kernel" / "Default" norm keys,
SURVEY.md §8b)."""
from __future__ import annotations


class Register(dict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._dict = {}

    def __call__(self, target_or_name):
        if callable(target_or_name):
            return self.register(target_or_name)
        return lambda x: self.register(x, key=target_or_name)

    def register(self, target, key=None):
        if not callable(target):
            raise Exception(f"Error: {target} must be callable!")
        if key is None:
            key = target.__name__
        if key in self._dict:
            raise Exception(f"{key} already exists.")
        self[key] = target
        return target

    def __setitem__(self, key, value):
        self._dict[key] = value

    def __getitem__(self, key):
        return self._dict[key]

    def __contains__(self, key):
        return key in self._dict

    def __str__(self):
        return str(self._dict)

    def keys(self):
        return self._dict.keys()

    def values(self):
        return self._dict.values()

    def items(self):
        return self._dict.items()


MM_WEIGHT_REGISTER = Register()
ATTN_WEIGHT_REGISTER = Register()
RMS_WEIGHT_REGISTER = Register()
LN_WEIGHT_REGISTER = Register()
TENSOR_REGISTER = Register()

# keys this package registers
MM_KEY = "B200-bf16"
ATTN_KEY = "b200_fmha"


def install_into_lightx2v() -> bool:
    """Register the B200 op classes into a LightX2V checkout that is importable in this process, so that a stock
    LightX2V config selects them with `mm_config.mm_type = "B200-bf16"` and `self_attn_1_type = cross_attn_1_type =
    cross_attn_2_type = "b200_fmha"`; the hard-coded norm keys are overridden in place (registry_factory.py:25-26).
    Returns False when LightX2V is not importable (stand-alone use of this package)."""
    try:
        from lightx2v.utils import registry_factory as rf  # type: ignore
    except Exception:
        return False
    from . import ops

    if MM_KEY not in rf.MM_WEIGHT_REGISTER:
        rf.MM_WEIGHT_REGISTER.register(ops.MMWeightB200, key=MM_KEY)
    for key, cls in ((ops.FP8_MM_KEY, ops.MMWeightFp8B200), (ops.NVFP4_MM_KEY, ops.MMWeightNvfp4B200)):
        if key not in rf.MM_WEIGHT_REGISTER:
            rf.MM_WEIGHT_REGISTER.register(cls, key=key)
    if ATTN_KEY not in rf.ATTN_WEIGHT_REGISTER:
        rf.ATTN_WEIGHT_REGISTER.register(ops.FmhaWeightB200, key=ATTN_KEY)
    rf.RMS_WEIGHT_REGISTER["sgl-kernel"] = ops.RMSWeightB200
    rf.LN_WEIGHT_REGISTER["Default"] = ops.LNWeightB200
    return True
