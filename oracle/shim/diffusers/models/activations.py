import torch.nn as nn
import torch.nn.functional as F

from .lora import LoRACompatibleLinear
from .._stubs import make_stub


def get_activation(name):
    name = name.lower()
    return {"silu": nn.SiLU, "swish": nn.SiLU, "mish": nn.Mish, "gelu": nn.GELU, "relu": nn.ReLU}[name]()


class GEGLU(nn.Module):
    """diffusers 0.25 GEGLU: proj to 2*dim_out, chunk into (value, gate), value * gelu(gate) (erf GELU)."""

    def __init__(self, dim_in, dim_out, bias=True):
        super().__init__()
        self.proj = LoRACompatibleLinear(dim_in, dim_out * 2, bias=bias)

    def forward(self, hidden_states, scale: float = 1.0):
        hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
        return hidden_states * F.gelu(gate)


GELU = make_stub("GELU")
ApproximateGELU = make_stub("ApproximateGELU")
