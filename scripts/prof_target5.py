"""ncu target (profiler range): short-K GEMMs with bias+residual epilogue at config-2 shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from idm_vton_b200 import lib as L  # noqa: E402

dev = "cuda"
L.load()


def rnd(*s, scale=1.0):
    return (torch.randn(*s, device=dev) * scale).half()


shapes = [(12288, 640, 640), (3072, 1280, 1280)]
ops = []
for (M, N, K) in shapes:
    a, w, b, r = rnd(M, K), rnd(N, K, scale=K ** -0.5), rnd(N), rnd(M, N)
    o = torch.empty(M, N, dtype=torch.float16, device=dev)
    ops.append((a, w, b, r, o))


def run():
    for (a, w, b, r, o) in ops:
        L.gemm(a, w, bias=b, residual=r, out=o)


for _ in range(3):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
