#!/bin/bash
mkdir -p gpurun_out
date > gpurun_out/attn.log
( timeout 420 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" --timeout 200 -p no:cacheprovider -x -s >> gpurun_out/attn.log 2>&1; echo "attention tests exit $?" | tee -a gpurun_out/attn.log; tail -n 8 gpurun_out/attn.log )
( timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/attn_timing.log | tail -n 12
import os, sys, json, torch
sys.path.insert(0, ".")
from idm_vton_b200 import lib as L
from scripts.microbench import timeit, timeit_graph, rnd
L.load()
for (B, H, N, Ng, tag) in [(4, 10, 3072, 3072, "L1 self+garment"), (4, 20, 768, 768, "L2 self+garment"), (16, 10, 3072, 0, "L1 garment self batch16"), (16, 20, 768, 0, "L2 garment self batch16")]:
    C = H * 64
    q, k, v = rnd(B, N, C), rnd(B, N, C), rnd(B, N, C)
    gk = rnd(max(B // 2, 1), Ng, C) if Ng else None
    gv = rnd(max(B // 2, 1), Ng, C) if Ng else None
    o = torch.empty(B, N, C, dtype=torch.float16, device='cuda')
    fl = 4.0 * B * H * N * N * 64 + (4.0 * (B // 2) * H * N * Ng * 64 if Ng else 0)
    res = {}
    for name, opts in (("attn6_poly0", {"attention_poly_exp": 0}), ("attn6_poly1", {"attention_poly_exp": 1}), ("attn6_poly2", {"attention_poly_exp": 2})):
        for kk, vv in opts.items():
            L.set_option(kk, vv)
        ms = timeit_graph(lambda: L.attention(q, k, v, gk, gv, kv1_off=B // 2, heads=H, out=o), n=10)
        res[name] = round(fl / ms / 1e9, 1)
    L.set_option("attention_16_warps", 1); L.set_option("attention_fp16_exp", 1); L.set_option("attention_poly_exp", 0)
    print(json.dumps({"tag": tag, "tflops": res}))
PY
)
( timeout 200 python - <<'PY' 2>&1 | tee gpurun_out/norm_timing.log | tail -n 6
import sys, json, torch
sys.path.insert(0, ".")
from idm_vton_b200 import lib as L
from scripts.microbench import timeit, rnd
L.load()
for (B, HW, C) in [(4, 3072, 640), (4, 768, 1280), (4, 12288, 320)]:
    xs, g, be = rnd(B, HW, C), rnd(C), rnd(C)
    gn = timeit(lambda: L.groupnorm(xs, g, be, 1e-5, True))
    ln = timeit(lambda: L.layernorm(xs.view(-1, C), g, be))
    mb = 2 * xs.numel() * 2 / 1e6
    print(json.dumps({"shape": [B, HW, C], "gn_us": round(gn * 1e3, 1), "gn_GBps": round(1.5 * mb / gn, 0), "ln_us": round(ln * 1e3, 1), "ln_GBps": round(mb / ln, 0)}))
PY
)
( timeout 280 python - <<'PY' 2>&1 | tee gpurun_out/vae_timing.log | tail -n 6
import sys, json, time, torch
sys.path.insert(0, ".")
from idm_vton_b200.vae import AutoencoderKL
torch.manual_seed(0)
vae = AutoencoderKL().cuda().float().eval()
x = torch.randn(2, 3, 1024, 768, device="cuda")
z = torch.randn(2, 4, 128, 96, device="cuda")
def t(fn, n=3):
    fn(); fn(); torch.cuda.synchronize(); t0 = time.time()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.time() - t0) / n * 1e3
with torch.no_grad():
    r = {"encode_autotune_ms": round(t(lambda: vae.encode(x)), 1), "decode_autotune_ms": round(t(lambda: vae.decode(z)), 1)}
    import idm_vton_b200.vae as V, contextlib
    V._conv_autotune = lambda x: contextlib.nullcontext()
    r.update({"encode_default_ms": round(t(lambda: vae.encode(x)), 1), "decode_default_ms": round(t(lambda: vae.decode(z)), 1)})
print(json.dumps(r))
PY
)
