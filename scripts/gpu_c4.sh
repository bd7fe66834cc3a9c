#!/bin/bash
mkdir -p gpurun_out
( timeout 240 python - <<'PY' 2>&1 | tee gpurun_out/c4.log | tail -n 40
import sys, json, torch
sys.path.insert(0, ".")
from idm_vton_b200 import lib as L
from idm_vton_b200.engine import pack_geglu
from scripts.microbench import timeit_graph, rnd
L.load()
def err(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1.0)).item()
# ---- correctness
for (M, N, K) in [(512, 512, 128), (3072, 1280, 1280), (1000, 768, 192), (3072, 3840, 1280), (256, 512, 64)]:
    a, w, b, r = rnd(M, K), rnd(N, K, scale=K ** -0.5), rnd(N), rnd(M, N)
    ref = (a.float() @ w.float().t())
    o = L.gemm(a, w, force_bn=2256)
    e0 = err(o, ref)
    o = L.gemm(a, w, bias=b, residual=r, force_bn=2256)
    ref2 = (ref + b.float()).half().float() + r.float()
    e1 = err(o, ref2)
    o1 = L.gemm(a, w, bias=b, residual=r, force_bn=1256)
    print(json.dumps({"check": [M, N, K], "plain_err": e0, "bias_res_err": e1, "same_as_np1": bool(torch.equal(o, o1))}), flush=True)
a, w, b = rnd(3072, 1280), rnd(10240, 1280, scale=1280 ** -0.5), rnd(10240)
wp, bp = pack_geglu(w, b, 256)
o2 = L.gemm(a, wp, bias=bp, geglu=True, force_bn=2256)
o1 = L.gemm(a, wp, bias=bp, geglu=True, force_bn=1256)
print(json.dumps({"check": "geglu", "same_as_np1": bool(torch.equal(o1, o2)), "max": o2.float().abs().max().item()}), flush=True)
# ---- timing
for (M, N, K, tag) in [(3072, 10240, 1280, "L2 ff1"), (3072, 1280, 5120, "L2 ff2"), (3072, 3840, 1280, "L2 qkv"), (3072, 1280, 1280, "L2 out"),
                       (12288, 5120, 640, "L1 ff1"), (24576, 1280, 1280, "garment L2 out b16"), (8192, 8192, 8192, "square 8k")]:
    a, w, b = rnd(M, K), rnd(N, K, scale=K ** -0.5), rnd(N)
    o = torch.empty(M, N, dtype=torch.float16, device="cuda")
    res = {}
    for name, f in (("np1", 1256), ("np2", 2256)):
        t = timeit_graph(lambda: L.gemm(a, w, bias=b, out=o, force_bn=f), n=10)
        res[name] = [round(1e3 * t, 1), round(2.0 * M * N * K / t / 1e9)]
    if "ff1" in tag:
        wp, bp = pack_geglu(w, b, 256)
        og = torch.empty(M, N // 2, dtype=torch.float16, device="cuda")
        for name, f in (("geglu_np1", 1256), ("geglu_np2", 2256)):
            t = timeit_graph(lambda: L.gemm(a, wp, bias=bp, geglu=True, out=og, force_bn=f), n=10)
            res[name] = [round(1e3 * t, 1), round(2.0 * M * N * K / t / 1e9)]
    print(json.dumps({"tag": tag, "us_tflops": res}), flush=True)
PY
)
