"""Short target for `ncu --set full`: a handful of representative launches of the dominant kernels at config-2 shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idm_vton_b200 import lib as L  # noqa: E402
from idm_vton_b200.engine import pack_conv3x3  # noqa: E402

dev = "cuda"
L.load()


def rnd(*s, scale=1.0):
    return (torch.randn(*s, device=dev) * scale).half()


a, w = rnd(3072, 1280), rnd(10240, 1280, scale=1280 ** -0.5)       # L2 FF1 (plain epilogue)
a1, w1 = rnd(12288, 640), rnd(1920, 640, scale=640 ** -0.5)        # L1 QKV
x = rnd(4, 64, 48, 640)
wc = pack_conv3x3(rnd(640, 640, 3, 3, scale=(9 * 640) ** -0.5))
bc = rnd(640)
q, k, v = rnd(4, 3072, 640), rnd(4, 3072, 640), rnd(4, 3072, 640)
gk, gv = rnd(2, 3072, 640), rnd(2, 3072, 640)
for _ in range(4):
    L.gemm(a, w, force_bn=256)
    L.gemm(a1, w1, force_bn=128)
    L.conv3x3(x, wc, bias=bc, force_bn=128)
    L.attention(q, k, v, gk, gv, kv1_off=2, heads=10)
torch.cuda.synchronize()
