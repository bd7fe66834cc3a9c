import logging as _pylogging
from collections import OrderedDict

import torch

USE_PEFT_BACKEND = False


class BaseOutput(OrderedDict):
    def __post_init__(self):
        for k, v in self.__dict__.items():
            self[k] = v

    def __getitem__(self, k):
        if isinstance(k, int):
            return list(self.values())[k]
        return super().__getitem__(k)


def deprecate(*args, **kwargs):
    return None


class logging:  # noqa: N801  (mirrors `diffusers.utils.logging`)
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)


def scale_lora_layers(*a, **k):
    return None


def unscale_lora_layers(*a, **k):
    return None


def is_torch_version(op, ver):
    from packaging import version
    cur = version.parse(torch.__version__.split("+")[0])
    ref = version.parse(ver)
    return {">=": cur >= ref, ">": cur > ref, "<": cur < ref, "<=": cur <= ref, "==": cur == ref}[op]


def is_invisible_watermark_available():
    return False


def is_torch_xla_available():
    return False


def replace_example_docstring(doc):
    def deco(fn):
        return fn
    return deco
