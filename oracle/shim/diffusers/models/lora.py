import torch.nn as nn
import torch.nn.functional as F

from .._stubs import stub_getattr


class LoRACompatibleLinear(nn.Linear):
    """diffusers 0.25 LoRACompatibleLinear without an attached lora_layer: plain nn.Linear, `scale` ignored."""

    def __init__(self, *args, lora_layer=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.lora_layer = lora_layer

    def forward(self, hidden_states, scale: float = 1.0):
        return F.linear(hidden_states, self.weight, self.bias)


class LoRACompatibleConv(nn.Conv2d):
    def __init__(self, *args, lora_layer=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.lora_layer = lora_layer

    def forward(self, hidden_states, scale: float = 1.0):
        return F.conv2d(hidden_states, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)


def adjust_lora_scale_text_encoder(*a, **k):
    return None


__getattr__ = stub_getattr(__name__)
