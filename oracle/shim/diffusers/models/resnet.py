"""Restated diffusers 0.25.0 ResnetBlock2D / Downsample2D / Upsample2D (SDXL code path only)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .activations import get_activation
from .lora import LoRACompatibleConv, LoRACompatibleLinear
from .._stubs import stub_getattr


class Upsample2D(nn.Module):
    def __init__(self, channels, use_conv=False, use_conv_transpose=False, out_channels=None, name="conv",
                 kernel_size=None, padding=1, norm_type=None, eps=None, elementwise_affine=None, bias=True,
                 interpolate=True):
        super().__init__()
        assert use_conv and not use_conv_transpose and norm_type is None
        self.channels = channels
        self.out_channels = out_channels or channels
        self.conv = LoRACompatibleConv(self.channels, self.out_channels, 3, padding=padding, bias=bias)

    def forward(self, hidden_states, output_size=None, scale: float = 1.0):
        dtype = hidden_states.dtype
        if dtype == torch.bfloat16:
            hidden_states = hidden_states.to(torch.float32)
        if hidden_states.shape[0] >= 64:
            hidden_states = hidden_states.contiguous()
        if output_size is None:
            hidden_states = F.interpolate(hidden_states, scale_factor=2.0, mode="nearest")
        else:
            hidden_states = F.interpolate(hidden_states, size=output_size, mode="nearest")
        if dtype == torch.bfloat16:
            hidden_states = hidden_states.to(dtype)
        return self.conv(hidden_states, scale)


class Downsample2D(nn.Module):
    def __init__(self, channels, use_conv=False, out_channels=None, padding=1, name="conv", kernel_size=3,
                 norm_type=None, eps=None, elementwise_affine=None, bias=True):
        super().__init__()
        assert use_conv and norm_type is None
        self.channels = channels
        self.out_channels = out_channels or channels
        self.padding = padding
        self.conv = LoRACompatibleConv(self.channels, self.out_channels, kernel_size, stride=2, padding=padding,
                                       bias=bias)

    def forward(self, hidden_states, scale: float = 1.0):
        return self.conv(hidden_states, scale)


class ResnetBlock2D(nn.Module):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=512,
                 groups=32, groups_out=None, pre_norm=True, eps=1e-6, non_linearity="swish", skip_time_act=False,
                 time_embedding_norm="default", kernel=None, output_scale_factor=1.0, use_in_shortcut=None, up=False,
                 down=False, conv_shortcut_bias=True, conv_2d_out_channels=None):
        super().__init__()
        assert time_embedding_norm == "default" and not up and not down
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.out_channels = out_channels
        self.output_scale_factor = output_scale_factor
        self.skip_time_act = skip_time_act
        groups_out = groups_out or groups
        self.norm1 = nn.GroupNorm(num_groups=groups, num_channels=in_channels, eps=eps, affine=True)
        self.conv1 = LoRACompatibleConv(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.time_emb_proj = LoRACompatibleLinear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(num_groups=groups_out, num_channels=out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(dropout)
        conv_2d_out_channels = conv_2d_out_channels or out_channels
        self.conv2 = LoRACompatibleConv(out_channels, conv_2d_out_channels, kernel_size=3, stride=1, padding=1)
        self.nonlinearity = get_activation(non_linearity)
        self.upsample = self.downsample = None
        self.use_in_shortcut = self.in_channels != conv_2d_out_channels if use_in_shortcut is None else use_in_shortcut
        self.conv_shortcut = None
        if self.use_in_shortcut:
            self.conv_shortcut = LoRACompatibleConv(in_channels, conv_2d_out_channels, kernel_size=1, stride=1,
                                                    padding=0, bias=conv_shortcut_bias)

    def forward(self, input_tensor, temb, scale: float = 1.0):
        hidden_states = input_tensor
        hidden_states = self.norm1(hidden_states)
        hidden_states = self.nonlinearity(hidden_states)
        hidden_states = self.conv1(hidden_states, scale)
        if self.time_emb_proj is not None:
            if not self.skip_time_act:
                temb = self.nonlinearity(temb)
            temb = self.time_emb_proj(temb, scale)[:, :, None, None]
        if temb is not None:
            hidden_states = hidden_states + temb
        hidden_states = self.norm2(hidden_states)
        hidden_states = self.nonlinearity(hidden_states)
        hidden_states = self.dropout(hidden_states)
        hidden_states = self.conv2(hidden_states, scale)
        if self.conv_shortcut is not None:
            input_tensor = self.conv_shortcut(input_tensor, scale)
        return (input_tensor + hidden_states) / self.output_scale_factor


__getattr__ = stub_getattr(__name__)
