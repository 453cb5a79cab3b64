"""Minimal nn.Module-like weight tree with the reference's interface
(lightx2v/common/modules/weight_module.py:1-182): add_module / load / to_cuda / to_cpu / state_dict / clear."""
from __future__ import annotations


class WeightModule:
    def __init__(self):
        self._modules = {}
        self._parameters = {}

    def add_module(self, name, module):
        self._modules[name] = module
        setattr(self, name, module)

    def register_parameter(self, name, param):
        self._parameters[name] = param
        setattr(self, name, param)

    def _children(self):
        yield from self._modules.values()
        yield from self._parameters.values()

    def load(self, weight_dict):
        mm_config = getattr(self, "config", {}).get("mm_config", None) if hasattr(self, "config") else None
        for child in self._children():
            if child is None:
                continue
            if hasattr(child, "set_config") and not isinstance(child, WeightModule):
                child.set_config(mm_config)
            if hasattr(child, "load"):
                child.load(weight_dict)

    def calculate_size(self):
        total = 0
        for child in self._children():
            if child is None:
                continue
            if isinstance(child, WeightModule):
                total += child.calculate_size()
            elif hasattr(child, "_calculate_size"):
                total += child._calculate_size()
        return total

    def clear(self):
        for child in self._children():
            if child is not None and hasattr(child, "clear"):
                child.clear()

    def state_dict(self, destination=None):
        if destination is None:
            destination = {}
        for child in self._children():
            if child is not None:
                child.state_dict(destination)
        return destination

    def to_cuda(self, non_blocking=False):
        for child in self._children():
            if child is not None and hasattr(child, "to_cuda"):
                child.to_cuda(non_blocking) if not isinstance(child, WeightModule) else child.to_cuda(non_blocking)

    def to_cpu(self, non_blocking=False):
        for child in self._children():
            if child is not None and hasattr(child, "to_cpu"):
                child.to_cpu(non_blocking)

    # the reference exposes *_async spellings used by the offload manager; on B200 nothing is offloaded, they alias
    def to_cuda_async(self):
        self.to_cuda(non_blocking=True)

    def to_cpu_async(self):
        self.to_cpu(non_blocking=True)


class WeightModuleList(WeightModule):
    def __init__(self, modules=None):
        super().__init__()
        self._list = []
        if modules is not None:
            for idx, module in enumerate(modules):
                self.append(module)

    def append(self, module):
        idx = len(self._list)
        self._list.append(module)
        self.add_module(str(idx), module)

    def __getitem__(self, idx):
        return self._list[idx]

    def __len__(self):
        return len(self._list)

    def __iter__(self):
        return iter(self._list)
