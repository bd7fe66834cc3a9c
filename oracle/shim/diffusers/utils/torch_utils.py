import torch


def maybe_allow_in_graph(cls):
    return cls


def apply_freeu(*a, **k):
    raise NotImplementedError


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """diffusers.utils.torch_utils.randn_tensor: draws on the generator's device, then moves."""
    layout = layout or torch.strided
    device = device or torch.device("cpu")
    rand_device = device
    if generator is not None:
        gen_device_type = generator.device.type if not isinstance(generator, list) else generator[0].device.type
        if gen_device_type != torch.device(device).type and gen_device_type == "cpu":
            rand_device = "cpu"
    return torch.randn(shape, generator=generator, device=rand_device, dtype=dtype, layout=layout).to(device)
