"""UNet executors of the denoising engine: weight pre-packing + the launch sequence of both SDXL UNets.

Host orchestration is Python; all math is in libb200vton.so (see lib.py). One denoise step is a fixed launch sequence
over static buffers, so it is captured once into a CUDA graph and replayed for every step.
"""
import torch


# ------------------------------------------------------------------------------------------------
# weight packing (load time, plain torch)
# ------------------------------------------------------------------------------------------------
def pack_conv3x3(w):
    """[Cout, Cin, 3, 3] -> [9, Cout, Cin] (tap = ky*3+kx), the layout b200vton_conv3x3_nhwc streams by TMA."""
    cout, cin = w.shape[0], w.shape[1]
    return w.permute(2, 3, 0, 1).reshape(9, cout, cin).contiguous()


def pad_channels(w_packed, cin_to=None, cout_to=None):
    """Zero-pad a packed conv weight [9, Cout, Cin] (conv_in: Cin 13 -> 64; conv_out: Cout 4 -> 16)."""
    nine, cout, cin = w_packed.shape
    cin_to = cin_to or cin
    cout_to = cout_to or cout
    out = torch.zeros((nine, cout_to, cin_to), dtype=w_packed.dtype, device=w_packed.device)
    out[:, :cout, :cin] = w_packed
    return out


def pack_conv3x3_s2(w):
    """[Cout, Cin, 3, 3] -> [Cout, 9*Cin] with K ordered tap-major, matching b200vton_im2col3x3_s2_nhwc."""
    cout, cin = w.shape[0], w.shape[1]
    return w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()


def pack_geglu(w, b, bn):
    """GEGLU proj weight [8C, C] (rows = [value 4C | gate 4C], diffusers GEGLU.chunk order) -> rows interleaved per
    output tile of width bn: [value bn/2 | gate bn/2], so one accumulator tile holds both halves of its channels."""
    n2 = w.shape[0]
    n = n2 // 2
    hb = bn // 2
    assert n % hb == 0
    wv, wg = w[:n].reshape(n // hb, hb, -1), w[n:].reshape(n // hb, hb, -1)
    wp = torch.cat([wv, wg], dim=1).reshape(n2, -1).contiguous()
    bp = None
    if b is not None:
        bv, bg = b[:n].reshape(n // hb, hb), b[n:].reshape(n // hb, hb)
        bp = torch.cat([bv, bg], dim=1).reshape(n2).contiguous()
    return wp, bp
