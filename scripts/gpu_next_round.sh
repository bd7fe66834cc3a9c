#!/bin/bash
# First GPU call of the next round: validate what was written after round 1's GPU budget ran out. Every step has its own
# tight timeout and writes to gpurun_out/, so a stall costs at most that step. ~6 GPU-minutes in total.
mkdir -p gpurun_out
L=gpurun_out/next_round.log
date > $L
step() { echo "=== $1" | tee -a $L; shift; ( "$@" ) >> $L 2>&1; echo "    exit $?" | tee -a $L; }
# 1. experimental kernels: fp32 NHWC GroupNorm, whole VAE through the NHWC route
step "experimental GPU tests" env B200VTON_EXPERIMENTAL=1 timeout 150 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "fp32_nhwc or vae_nhwc" --timeout 100 -p no:cacheprovider
# 2. VAE timing: default route vs NHWC route (per call: encode B=2 at 1024x768, decode B=2)
step "VAE timing default" timeout 120 python - <<'PY'
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
print(json.dumps(r))
PY
# 2b. deep-pipeline GEMM variant: parity, then A/B timing on the big 256-wide GEMMs (CUDA-graph timed)
step "deep-pipeline GEMM test" env B200VTON_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "deep_pipeline" --timeout 100 -p no:cacheprovider
step "deep-pipeline GEMM timing" timeout 150 python - <<'PY'
import sys, json, torch
sys.path.insert(0, ".")
from idm_vton_b200 import lib as L
from idm_vton_b200.engine import pack_geglu
from scripts.microbench import timeit_graph, rnd
L.load()
for (M, N, K, tag) in [(3072, 10240, 1280, "L2 ff1"), (3072, 1280, 5120, "L2 ff2"), (3072, 3840, 1280, "L2 qkv"), (24576, 1280, 1280, "garment out b16"), (8192, 8192, 8192, "square 8k")]:
    a, w, b = rnd(M, K), rnd(N, K, scale=K ** -0.5), rnd(N)
    o = torch.empty(M, N, dtype=torch.float16, device="cuda")
    res = {}
    for deep in (0, 1):
        L.set_option("gemm_deep_pipeline", deep)
        t = timeit_graph(lambda: L.gemm(a, w, bias=b, out=o, force_bn=1256), n=10)
        res[f"deep{deep}"] = [round(1e3 * t, 1), round(2.0 * M * N * K / t / 1e9)]
        if "ff1" in tag:
            wp, bp = pack_geglu(w, b, 256)
            og = torch.empty(M, N // 2, dtype=torch.float16, device="cuda")
            t = timeit_graph(lambda: L.gemm(a, wp, bias=bp, geglu=True, out=og, force_bn=1256), n=10)
            res[f"geglu_deep{deep}"] = [round(1e3 * t, 1), round(2.0 * M * N * K / t / 1e9)]
    L.set_option("gemm_deep_pipeline", 0)
    print(json.dumps({"tag": tag, "us_tflops": res}))
PY
# 3. programmatic dependent launch with the wait-before-alloc fix: the e2e section is where it stalled (2 of 5 runs)
for i in 1 2 3 4; do
  step "bench with PDL on, run $i" env B200VTON_PDL=1 B200VTON_E2E_TIMEOUT=120 timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline
done
grep -h '"metric"' $L | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('PDL run:', round(d['value'], 3), d['e2e'])" | tee -a $L
tail -n 30 $L
