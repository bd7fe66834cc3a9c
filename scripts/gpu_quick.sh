#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" --timeout 120 -p no:cacheprovider -x > gpurun_out/quick.log 2>&1; echo "exit $?" >> gpurun_out/quick.log; tail -n 6 gpurun_out/quick.log )
( timeout 400 python -m pytest tests/test_engine_gpu.py -q -m gpu --timeout 200 -p no:cacheprovider -x > gpurun_out/quick_engine.log 2>&1; echo "exit $?" >> gpurun_out/quick_engine.log; tail -n 6 gpurun_out/quick_engine.log )
( timeout 400 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; tail -c 1500 gpurun_out/quick_bench.json )
( timeout 200 python - <<'PY' 2>&1 | tee gpurun_out/xattn_timing.log | tail -n 6
import sys, json, torch
sys.path.insert(0, ".")
from idm_vton_b200 import lib as L
from scripts.microbench import timeit, rnd
L.load()
for (B, H, N, tag) in [(4, 20, 768, "L2"), (4, 10, 3072, "L1")]:
    C = H * 64
    q = rnd(B, N, C)
    kt, vt, ki, vi = rnd(B, 77, C), rnd(B, 77, C), rnd(B, 16, C), rnd(B, 16, C)
    out = torch.empty_like(q)
    t_text = timeit(lambda: L.attention(q, kt, vt, heads=H, out=out))
    t_ip = timeit(lambda: L.attention(q, ki, vi, heads=H, out=out, accumulate=True))
    t_f = timeit(lambda: L.cross_attention(q, kt, vt, ki, vi, heads=H, out=out))
    t_t = timeit(lambda: L.cross_attention(q, kt, vt, heads=H, out=out))
    print(json.dumps({"tag": tag, "text_us": round(t_text * 1e3, 1), "ip_us": round(t_ip * 1e3, 1), "fused_us": round(t_f * 1e3, 1), "fused_text_only_us": round(t_t * 1e3, 1)}))
PY
)
