#!/bin/bash
mkdir -p gpurun_out
( timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_r1_v6 -f \
    python scripts/prof_target5.py > gpurun_out/ncu_full6.log 2>&1; echo "ncu full exit $?"; tail -n 2 gpurun_out/ncu_full6.log; ls -la gpurun_out/prof_r1_v6.ncu-rep )
( timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/gemm_variants.log | tail -n 30
import sys, json, torch
sys.path.insert(0, ".")
from idm_vton_b200 import lib as L
from scripts.microbench import timeit_graph, rnd
L.load()
for (M, N, K, tag) in [(12288, 640, 640, "L1 out"), (3072, 1280, 1280, "L2 out"), (12288, 1920, 640, "L1 qkv")]:
    a, w, b, r = rnd(M, K), rnd(N, K, scale=K ** -0.5), rnd(N), rnd(M, N)
    o = torch.empty(M, N, dtype=torch.float16, device="cuda")
    res = {}
    for name, kw in (("auto_bias_res", dict(bias=b, residual=r)), ("auto_plain", dict()), ("v2_128_bias_res", dict(bias=b, residual=r, force_bn=1128)),
                     ("v2_128_plain", dict(force_bn=1128)), ("v1_128_bias_res", dict(bias=b, residual=r, force_bn=128)), ("v2_256_plain", dict(force_bn=1256))):
        if "256" in name and N % 256: continue
        try:
            t = timeit_graph(lambda: L.gemm(a, w, out=o, **kw))
            res[name] = [round(1e3 * t, 1), round(2.0 * M * N * K / t / 1e9)]
        except Exception as e:
            res[name] = str(e)[:60]
    print(json.dumps({"tag": tag, "us_tflops": res}))
PY
)
