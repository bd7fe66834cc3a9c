#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "fp32_tf32" --timeout 100 -p no:cacheprovider -x -s > gpurun_out/vae_conv_test.log 2>&1; echo "exit $?" >> gpurun_out/vae_conv_test.log; tail -n 12 gpurun_out/vae_conv_test.log )
( timeout 280 python - <<'PY' 2>&1 | tee gpurun_out/vae_timing.log | tail -n 8
import sys, json, time, torch
sys.path.insert(0, ".")
from idm_vton_b200.vae import AutoencoderKL
import idm_vton_b200.vae as V
from idm_vton_b200 import lib as L
from scripts.microbench import timeit
torch.manual_seed(0)
vae = AutoencoderKL().cuda().float().eval()
x = torch.randn(2, 3, 1024, 768, device="cuda")
z = torch.randn(2, 4, 128, 96, device="cuda")
def t(fn, n=3):
    fn(); fn(); torch.cuda.synchronize(); t0 = time.time()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.time() - t0) / n * 1e3
with torch.no_grad():
    e1 = vae.encode(x).latent_dist.mean; d1 = vae.decode(z).sample
    r = {"encode_engine_ms": round(t(lambda: vae.encode(x)), 1), "decode_engine_ms": round(t(lambda: vae.decode(z)), 1)}
    orig = V._conv
    V._conv = lambda conv, x: conv(x)
    e0 = vae.encode(x).latent_dist.mean; d0 = vae.decode(z).sample
    r.update({"encode_cudnn_ms": round(t(lambda: vae.encode(x)), 1), "decode_cudnn_ms": round(t(lambda: vae.decode(z)), 1)})
    r["encode_maxdiff_rel"] = ((e1 - e0).abs().max() / e0.abs().max()).item()
    r["decode_maxdiff_rel"] = ((d1 - d0).abs().max() / d0.abs().max()).item()
    V._conv = orig
print(json.dumps(r))
for (B, C, Co, H, W) in [(2, 128, 128, 1024, 768), (2, 256, 256, 512, 384), (2, 512, 512, 256, 192), (2, 512, 512, 128, 96)]:
    xx = torch.randn(B, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    w = torch.randn(Co, C, 3, 3, device="cuda") * 0.01
    wp = L.pack_conv3x3_f32(w)
    ms = timeit(lambda: L.conv3x3_f32(xx, wp, None), iters=5)
    xc = xx.contiguous()
    ms2 = timeit(lambda: torch.nn.functional.conv2d(xc, w, None, padding=1), iters=5)
    fl = 2.0 * B * H * W * 9 * C * Co
    print(json.dumps({"conv": [B, C, Co, H, W], "engine_ms": round(ms, 3), "engine_tflops": round(fl / ms / 1e9, 1), "cudnn_ms": round(ms2, 3), "cudnn_tflops": round(fl / ms2 / 1e9, 1)}))
PY
)
