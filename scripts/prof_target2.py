"""ncu target: ping-pong attention and the 2-CTA GEMM at config-2 shapes (a few launches each)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idm_vton_b200 import lib as L  # noqa: E402

dev = "cuda"
L.load()


def rnd(*s, scale=1.0):
    return (torch.randn(*s, device=dev) * scale).half()


q, k, v = rnd(4, 3072, 640), rnd(4, 3072, 640), rnd(4, 3072, 640)
gk, gv = rnd(2, 3072, 640), rnd(2, 3072, 640)
a, w = rnd(12288, 640), rnd(5120, 640, scale=640 ** -0.5)
a2, w2 = rnd(3072, 1280), rnd(1280, 1280, scale=1280 ** -0.5)
for _ in range(3):
    L.attention(q, k, v, gk, gv, kv1_off=2, heads=10)
    L.gemm(a, w, force_bn=1256)
    L.gemm(a2, w2, force_bn=1256)
torch.cuda.synchronize()
