#!/bin/bash
# Round 2, GPU call 3: graph-timed small kernels, launch list of one try-on step, ncu --set full of the dominant kernels,
# PDL-in-graph A/B, VAE op breakdown.
mkdir -p gpurun_out
L=gpurun_out/r2_call3.log
date > $L
step() { echo "=== $1" | tee -a $L; shift; ( "$@" ) >> $L 2>&1; echo "    exit $?" | tee -a $L; }
step "memcheck on the test that crashed in call 2" timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_engine_gpu.py -q -m gpu -k "tiny_unets_vs_oracle and 2-16-24" -p no:cacheprovider --timeout 400
for f in tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_seams_gpu.py tests/test_fullsize_gpu.py; do
  step "pytest $f (own process)" timeout 900 python -m pytest $f -q -m gpu -s --timeout 600 -p no:cacheprovider
done
step "graph-timed small kernels" timeout 300 python - <<'PY'
import sys, json, torch
sys.path.insert(0, ".")
from idm_vton_b200 import lib as L
from scripts.microbench import timeit_graph, rnd
L.load()
def p(**kw): print(json.dumps(kw), flush=True)
for (B, HW, C, C1) in [(4, 3072, 640, 0), (4, 768, 1280, 0), (4, 12288, 320, 0), (4, 768, 2560, 0), (4, 3072, 1280, 0), (4, 12288, 640, 320), (16, 12288, 320, 0), (2, 12288, 320, 0)]:
    xs, g, be = rnd(B, HW, C), rnd(C + C1), rnd(C + C1)
    x1 = rnd(B, HW, C1) if C1 else None
    t = timeit_graph(lambda: L.groupnorm(xs, g, be, 1e-5, True, x1=x1))
    mb = B * HW * (C + C1) * 4 / 1e6
    p(op="groupnorm_one_launch", shape=[B, HW, C + C1], us=round(1e3 * t, 2), algorithmic_mbytes=round(mb, 1), gbs=round(mb / t / 1e3 * 1e3 / 1e3, 1))
for (rows, C) in [(12288, 640), (3072, 1280)]:
    xs, g, be = rnd(rows, C), rnd(C), rnd(C)
    p(op="layernorm", shape=[rows, C], us=round(1e3 * timeit_graph(lambda: L.layernorm(xs, g, be)), 2))
for (M, N, K, tag) in [(3072, 1280, 1280, "L2 out/q2"), (3072, 3840, 1280, "L2 qkv"), (3072, 1280, 5120, "L2 ff2"), (12288, 640, 640, "L1 out"), (12288, 1920, 640, "L1 qkv"), (12288, 640, 2560, "L1 ff2"), (3072, 10240, 1280, "L2 ff1 plain")]:
    a, w, b, r = rnd(M, K), rnd(N, K, scale=K ** -0.5), rnd(N), rnd(M, N)
    t = timeit_graph(lambda: L.gemm(a, w, bias=b, residual=r))
    p(op="gemm_bias_res_auto", tag=tag, shape=[M, N, K], us=round(1e3 * t, 2), tflops=round(2.0 * M * N * K / t / 1e9, 1))
from idm_vton_b200.engine import pack_conv3x3
for (B, H, W, C) in [(4, 128, 96, 320), (4, 64, 48, 640)]:
    x, wp, b = rnd(B, H, W, C), pack_conv3x3(rnd(C, C, 3, 3, scale=(9 * C) ** -0.5)), rnd(C)
    t = timeit_graph(lambda: L.conv3x3(x, wp, bias=b, stride=2), n=10)
    p(op="downsample_conv_s2", shape=[B, H, W, C], us=round(1e3 * t, 2), tflops=round(2.0 * B * (H // 2) * (W // 2) * C * C * 9 / t / 1e9, 1))
for deep in (0, 1):
    L.set_option("gemm_deep_pipeline", deep)
    for (M, N, K, tag) in [(3072, 10240, 1280, "L2 ff1 plain"), (3072, 1280, 5120, "L2 ff2"), (8192, 8192, 8192, "square 8k")]:
        a, w, b = rnd(M, K), rnd(N, K, scale=K ** -0.5), rnd(N)
        o = torch.empty(M, N, dtype=torch.float16, device="cuda")
        t = timeit_graph(lambda: L.gemm(a, w, bias=b, out=o, force_bn=1256), n=10)
        p(op="gemm_deep_pipeline_ab", deep=deep, tag=tag, us=round(1e3 * t, 1), tflops=round(2.0 * M * N * K / t / 1e9))
L.set_option("gemm_deep_pipeline", 0)
PY
( timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_tryon_step.csv \
    python bench.py --profile-one-step --no-e2e --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_ncu_launchlist.log 2>&1; echo "ncu launch list exit $?" | tee -a $L; wc -l gpurun_out/r2_launches_tryon_step.csv | tee -a $L )
( timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_r2 -f \
    python scripts/prof_target_r2.py > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full exit $?" | tee -a $L; ls -la gpurun_out/prof_r2.ncu-rep | tee -a $L )
for pdl in 0 1; do
  echo "=== bench loop only, PDL in graph = $pdl" | tee -a $L
  B200VTON_PDL_GRAPH=$pdl timeout 400 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_pdlgraph_$pdl.json 2> gpurun_out/r2_bench_pdlgraph_$pdl.err; echo "    exit $?" | tee -a $L
done
for i in 1 2 3; do
  echo "=== bench with e2e, PDL in graph = 1, run $i" | tee -a $L
  B200VTON_PDL_GRAPH=1 B200VTON_E2E_TIMEOUT=150 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_pdlgraph_e2e_$i.json 2> gpurun_out/r2_bench_pdlgraph_e2e_$i.err; echo "    exit $?" | tee -a $L
done
step "VAE op breakdown (NHWC route)" timeout 200 python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
from idm_vton_b200.vae import AutoencoderKL
vae = AutoencoderKL().cuda().float().eval()
x, z = torch.randn(2, 3, 1024, 768, device="cuda"), torch.randn(2, 4, 128, 96, device="cuda")
from torch.profiler import profile, ProfilerActivity
with torch.no_grad():
    for _ in range(2): vae.encode(x); vae.decode(z)
    torch.cuda.synchronize()
    for name, fn in (("encode", lambda: vae.encode(x)), ("decode", lambda: vae.decode(z))):
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn(); torch.cuda.synchronize()
        print("=====", name)
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
PY
grep -h '"metric"' gpurun_out/r2_bench_pdlgraph_*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('BENCH value', round(d['value'], 4), 'ms/loop', round(d['ms_per_step'],1), 'e2e', d['e2e'] and d['e2e'].get('value'))" | tee -a $L
tail -n 120 $L
