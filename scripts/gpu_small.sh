#!/bin/bash
# graph-timed small kernels (no host time) + serialized launch list of one try-on denoise step
mkdir -p gpurun_out
( timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/small_timing.log | tail -n 30
import sys, json, torch
sys.path.insert(0, ".")
from idm_vton_b200 import lib as L
from scripts.microbench import timeit_graph, rnd
L.load()
def p(**kw): print(json.dumps(kw), flush=True)
for (B, H, N, tag) in [(4, 20, 768, "L2"), (4, 10, 3072, "L1")]:
    C = H * 64
    q = rnd(B, N, C); kt, vt, ki, vi = rnd(B, 77, C), rnd(B, 77, C), rnd(B, 16, C), rnd(B, 16, C)
    out = torch.empty_like(q)
    k, v = rnd(B, N, C), rnd(B, N, C); gk, gv = rnd(B // 2, N, C), rnd(B // 2, N, C)
    p(op="cross_fused", tag=tag, us=round(1e3 * timeit_graph(lambda: L.cross_attention(q, kt, vt, ki, vi, heads=H, out=out)), 2))
    p(op="cross_two_launch", tag=tag, us=round(1e3 * timeit_graph(lambda: (L.attention(q, kt, vt, heads=H, out=out), L.attention(q, ki, vi, heads=H, out=out, accumulate=True))), 2))
    t = timeit_graph(lambda: L.attention(q, k, v, gk, gv, kv1_off=B // 2, heads=H, out=out))
    p(op="attn1", tag=tag, us=round(1e3 * t, 2), tflops=round((4.0 * B * H * N * N * 64 * 1.5) / t / 1e9, 1))
for (B, HW, C) in [(4, 3072, 640), (4, 768, 1280), (4, 12288, 320), (4, 768, 2560), (4, 3072, 1280)]:
    xs, g, be = rnd(B, HW, C), rnd(C), rnd(C)
    p(op="groupnorm", shape=[B, HW, C], us=round(1e3 * timeit_graph(lambda: L.groupnorm(xs, g, be, 1e-5, True)), 2), mbytes=round(xs.numel() * 4 / 1e6, 1))
for (rows, C) in [(12288, 640), (3072, 1280)]:
    xs, g, be = rnd(rows, C), rnd(C), rnd(C)
    p(op="layernorm", shape=[rows, C], us=round(1e3 * timeit_graph(lambda: L.layernorm(xs, g, be)), 2), mbytes=round(xs.numel() * 4 / 1e6, 1))
for (M, N, K, tag) in [(3072, 1280, 1280, "L2 out/q2"), (3072, 3840, 1280, "L2 qkv"), (3072, 1280, 5120, "L2 ff2"), (12288, 640, 640, "L1 out"), (12288, 1920, 640, "L1 qkv"), (12288, 640, 2560, "L1 ff2"), (3072, 10240, 1280, "L2 ff1 plain")]:
    a, w, b, r = rnd(M, K), rnd(N, K, scale=K ** -0.5), rnd(N), rnd(M, N)
    t = timeit_graph(lambda: L.gemm(a, w, bias=b, residual=r))
    p(op="gemm_bias_res", tag=tag, shape=[M, N, K], us=round(1e3 * t, 2), tflops=round(2.0 * M * N * K / t / 1e9, 1))
PY
)
( timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v6.csv \
    python bench.py --profile-one-step --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench6.log 2>&1; echo "ncu launch list exit $?"; wc -l gpurun_out/launches_v6.csv )
