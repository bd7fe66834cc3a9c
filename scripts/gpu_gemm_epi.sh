#!/bin/bash
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 200 -p no:cacheprovider -x -k "gemm or conv or linear" > gpurun_out/epi_kernels.log 2>&1; echo "exit $?" >> gpurun_out/epi_kernels.log; tail -n 4 gpurun_out/epi_kernels.log )
( timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/gemm_variants2.log | tail -n 30
import sys, json, torch
sys.path.insert(0, ".")
from idm_vton_b200 import lib as L
from scripts.microbench import timeit_graph, rnd
L.load()
for (M, N, K, tag) in [(12288, 640, 640, "L1 out"), (3072, 1280, 1280, "L2 out"), (12288, 640, 2560, "L1 ff2"), (3072, 1280, 5120, "L2 ff2")]:
    a, w, b, r = rnd(M, K), rnd(N, K, scale=K ** -0.5), rnd(N), rnd(M, N)
    o = torch.empty(M, N, dtype=torch.float16, device="cuda")
    res = {}
    for name, kw in (("auto_bias_res", dict(bias=b, residual=r)), ("auto_bias", dict(bias=b)), ("auto_plain", dict())):
        t = timeit_graph(lambda: L.gemm(a, w, out=o, **kw))
        res[name] = [round(1e3 * t, 1), round(2.0 * M * N * K / t / 1e9)]
    print(json.dumps({"tag": tag, "us_tflops": res}))
PY
)
( timeout 400 python -m pytest tests/test_engine_gpu.py -q -m gpu --timeout 200 -p no:cacheprovider -x > gpurun_out/epi_engine.log 2>&1; echo "exit $?" >> gpurun_out/epi_engine.log; tail -n 3 gpurun_out/epi_engine.log )
( timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/epi_bench.json 2> gpurun_out/epi_bench.err; python -c "import json;d=json.load(open('gpurun_out/epi_bench.json'));print('bench', d['value'], d['ms_per_step'])" )
