"""String-keyed op registries with the reference's semantics (lightx2v/utils/registry_factory.py:1-56):
`@REGISTER("key")` decorates a class, duplicate keys raise, `REGISTER[key]` looks a class up, `REGISTER[key] = cls`
overrides without a check (that asymmetry is what lets a plugin replace the hard-coded "sgl-kernel" / "Default" norm keys,
SURVEY.md §8b)."""
from __future__ import annotations


class Register:
    """Stand-alone key -> class table for this package (when LightX2V itself is importable its own registries are used instead, see
    `install_into_lightx2v`).  Contract relied on by the weight trees and by plugins: decorating with a key that is already taken is an
    error, while plain item assignment replaces an entry."""

    __slots__ = ("_table",)

    def __init__(self):
        self._table = {}

    def __call__(self, key_or_class):
        if isinstance(key_or_class, str):
            return lambda cls: self.register(cls, key=key_or_class)
        return self.register(key_or_class)

    def register(self, cls, key=None):
        key = getattr(cls, "__name__", None) if key is None else key
        if not callable(cls) or key is None:
            raise TypeError(f"cannot register {cls!r}")
        if key in self._table:
            raise KeyError(f"{key} already exists.")
        self._table[key] = cls
        return cls

    def __setitem__(self, key, cls):
        self._table[key] = cls

    def __getitem__(self, key):
        return self._table[key]

    def __contains__(self, key):
        return key in self._table

    def __iter__(self):
        return iter(self._table)

    def __len__(self):
        return len(self._table)

    def keys(self):
        return self._table.keys()

    def items(self):
        return self._table.items()

    def __repr__(self):
        return f"Register({sorted(self._table)})"


MM_WEIGHT_REGISTER = Register()
ATTN_WEIGHT_REGISTER = Register()
RMS_WEIGHT_REGISTER = Register()
LN_WEIGHT_REGISTER = Register()
TENSOR_REGISTER = Register()

# keys this package registers
MM_KEY = "B200-bf16"
ATTN_KEY = "b200_fmha"


_saved_reference_entries = {}


def install_into_lightx2v() -> bool:
    """Register the B200 op classes into a LightX2V checkout that is importable in this process, so that a stock
    LightX2V config selects them with `mm_config.mm_type = "B200-bf16"` and `self_attn_1_type = cross_attn_1_type =
    cross_attn_2_type = "b200_fmha"`; the hard-coded norm keys are overridden in place (registry_factory.py:25-26; the entries they
    replace are remembered for `uninstall_from_lightx2v`).  Returns False when LightX2V is not importable (stand-alone use of this package)."""
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
    if MM_KEY not in rf.LN_WEIGHT_REGISTER:          # CogVideoX indexes the LayerNorm registry with the mm_type string
        rf.LN_WEIGHT_REGISTER.register(ops.LNWeightB200, key=MM_KEY)
    for reg_name, key, cls in (("RMS_WEIGHT_REGISTER", "sgl-kernel", ops.RMSWeightB200), ("LN_WEIGHT_REGISTER", "Default", ops.LNWeightB200)):
        reg = getattr(rf, reg_name)
        if key in reg and reg[key] is not cls:
            _saved_reference_entries[(reg_name, key)] = reg[key]
        reg[key] = cls
    return True


def uninstall_from_lightx2v() -> None:
    """Put the reference's own norm classes back under the keys `install_into_lightx2v` overrode (the added keys stay registered: they do
    not shadow anything)."""
    try:
        from lightx2v.utils import registry_factory as rf  # type: ignore
    except Exception:
        return
    for (reg_name, key), cls in list(_saved_reference_entries.items()):
        getattr(rf, reg_name)[key] = cls
        del _saved_reference_entries[(reg_name, key)]
