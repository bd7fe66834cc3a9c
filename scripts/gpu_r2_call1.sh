#!/bin/bash
# Round 2, GPU call 1: full-size integrated parity on the production kernels, the upgraded smoke(), validation of the
# experimental paths written at the end of round 1 (fp32 NHWC GroupNorm / VAE route, deep-pipeline GEMM), baseline bench.
mkdir -p gpurun_out
L=gpurun_out/r2_call1.log
date > $L
rm -f gpurun_out/fullsize_parity.jsonl
step() { echo "=== $1" | tee -a $L; shift; ( "$@" ) >> $L 2>&1; echo "    exit $?" | tee -a $L; }
step "full-size parity" timeout 600 python -m pytest tests/test_fullsize_gpu.py -x -q -m gpu -s --timeout 500 -p no:cacheprovider
step "smoke" timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
step "experimental GPU tests" env B200VTON_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "fp32_nhwc or vae_nhwc or deep_pipeline" --timeout 150 -p no:cacheprovider
step "VAE timing" timeout 150 python - <<'PY'
import sys, time, json, torch
sys.path.insert(0, ".")
from idm_vton_b200.vae import AutoencoderKL
import idm_vton_b200.vae as V
vae = AutoencoderKL().cuda().float().eval()
x, z = torch.randn(2, 3, 1024, 768, device="cuda"), torch.randn(2, 4, 128, 96, device="cuda")
def t(fn, n=3):
    fn(); fn(); torch.cuda.synchronize(); t0 = time.time()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return round((time.time() - t0) / n * 1e3, 1)
with torch.no_grad():
    r = {"encode_default_ms": t(lambda: vae.encode(x)), "decode_default_ms": t(lambda: vae.decode(z))}
    V._ENGINE_NHWC = True
    r.update({"encode_nhwc_ms": t(lambda: vae.encode(x)), "decode_nhwc_ms": t(lambda: vae.decode(z))})
    vh = AutoencoderKL().cuda().half().eval()
    V._ENGINE_NHWC = False
    xh, zh = x.half(), z.half()
    r.update({"encode_fp16_cudnn_ms": t(lambda: vh.encode(xh)), "decode_fp16_cudnn_ms": t(lambda: vh.decode(zh))})
    vb = AutoencoderKL().cuda().bfloat16().eval()
    xb, zb = x.bfloat16(), z.bfloat16()
    r.update({"encode_bf16_cudnn_ms": t(lambda: vb.encode(xb)), "decode_bf16_cudnn_ms": t(lambda: vb.decode(zb))})
print(json.dumps(r))
PY
step "bench baseline" env B200VTON_TRACE=1 B200VTON_E2E_TIMEOUT=200 timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline
tail -n 60 $L
