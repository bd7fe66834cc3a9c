"""Stub factory for the diffusers shim: names that the reference imports but never executes on the SDXL inference
path resolve to empty placeholder classes (module-level __getattr__, PEP 562)."""
import torch.nn as nn


def make_stub(name):
    def __init__(self, *a, **k):
        raise NotImplementedError(f"diffusers shim: {name} is a placeholder (not on the IDM-VTON hot path)")
    return type(name, (nn.Module,), {"__init__": __init__, "__module__": "diffusers._stubs"})


def stub_getattr(modname):
    cache = {}

    def __getattr__(name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name not in cache:
            cache[name] = make_stub(name)
        return cache[name]
    return __getattr__
