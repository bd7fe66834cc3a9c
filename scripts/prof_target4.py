"""ncu target (profiler range): one launch each of the dominant GEGLU GEMM, the P-in-TMEM attention, LN and GN at
config-2 shapes, after warm-up, with an L2 flush before the range."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idm_vton_b200 import lib as L  # noqa: E402
from idm_vton_b200.engine import pack_geglu  # noqa: E402

dev = "cuda"
L.load()


def rnd(*s, scale=1.0):
    return (torch.randn(*s, device=dev) * scale).half()


a, w, b = rnd(3072, 1280), rnd(10240, 1280, scale=1280 ** -0.5), rnd(10240)     # L2 FF1 (GEGLU), the dominant kernel
wp, bp = pack_geglu(w, b, 256)
q, k, v = rnd(4, 3072, 640), rnd(4, 3072, 640), rnd(4, 3072, 640)
gk, gv = rnd(2, 3072, 640), rnd(2, 3072, 640)
xs, g, be = rnd(4, 3072, 640), rnd(640), rnd(640)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def run():
    L.gemm(a, wp, bias=bp, geglu=True)
    L.attention(q, k, v, gk, gv, kv1_off=2, heads=10)
    L.groupnorm(xs, g, be, 1e-5, True)
    L.layernorm(xs.view(-1, 640), g, be)


for _ in range(3):
    run()
flush.fill_(1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
