"""Minimal `DiffusionPipeline` for the shim: exactly what src/tryon_pipeline.py uses of it — `register_modules`,
`register_to_config` / `.config`, `_execution_device`, `progress_bar`, `maybe_free_model_hooks`, `.to`."""
import contextlib

import torch

from ..configuration_utils import ConfigMixin


class _Bar:
    def update(self, n=1):
        pass


class DiffusionPipeline(ConfigMixin):
    def __init__(self):
        object.__setattr__(self, "_internal_dict", {})
        self._module_names = []

    def register_modules(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)
            self._module_names.append(k)

    @property
    def _execution_device(self):
        return next(self.unet.parameters()).device

    @contextlib.contextmanager
    def progress_bar(self, iterable=None, total=None):
        yield _Bar()

    def maybe_free_model_hooks(self):
        pass

    def to(self, *args, **kwargs):
        for k in self._module_names:
            m = getattr(self, k)
            if isinstance(m, torch.nn.Module):
                m.to(*args, **kwargs)
        return self
