#!/bin/bash
# Round 2, GPU call 4: pinpoint the state-dependent crash of tests/test_engine_gpu.py (passes alone and under memcheck,
# fails after two other tests of the same process), re-run the fixed seam tests and the re-pipelined GroupNorm.
mkdir -p gpurun_out
L=gpurun_out/r2_call4.log
date > $L
step() { echo "=== $1" | tee -a $L; shift; ( "$@" ) >> $L 2>&1; echo "    exit $?" | tee -a $L; }
step "engine tests, CUDA_LAUNCH_BLOCKING=1, stop at first failure" env CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests/test_engine_gpu.py -x -q -m gpu -s --timeout 500 -p no:cacheprovider
step "engine tests (first four) under memcheck in ONE process" timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_engine_gpu.py -q -m gpu -k "golden or tiny_unets" -p no:cacheprovider --timeout 800
step "engine tests plain" timeout 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -s --timeout 500 -p no:cacheprovider
step "seam tests" timeout 600 python -m pytest tests/test_seams_gpu.py -q -m gpu -s --timeout 500 -p no:cacheprovider
step "kernel tests" timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 500 -p no:cacheprovider
step "graph-timed GroupNorm" timeout 300 python - <<'PY'
import sys, json, torch
sys.path.insert(0, ".")
from idm_vton_b200 import lib as L
from scripts.microbench import timeit_graph, rnd
L.load()
for (B, HW, C, C1) in [(4, 3072, 640, 0), (4, 768, 1280, 0), (4, 12288, 320, 0), (4, 768, 2560, 0), (4, 3072, 1280, 0), (4, 12288, 640, 320), (16, 12288, 320, 0), (2, 12288, 320, 0)]:
    xs, g, be = rnd(B, HW, C), rnd(C + C1), rnd(C + C1)
    x1 = rnd(B, HW, C1) if C1 else None
    for silu in (True, False):
        t = timeit_graph(lambda: L.groupnorm(xs, g, be, 1e-5, silu, x1=x1))
        mb = B * HW * (C + C1) * 4 / 1e6
        print(json.dumps(dict(op="groupnorm_one_launch_v2", shape=[B, HW, C + C1], silu=silu, us=round(1e3 * t, 2), algorithmic_mbytes=round(mb, 1), tbs=round(mb / t / 1e6, 2))), flush=True)
PY
echo "=== bench default" | tee -a $L
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_call4.json 2> gpurun_out/r2_bench_call4.err; echo "    exit $?" | tee -a $L
tail -n 3 gpurun_out/r2_bench_call4.err >> $L
tail -n 150 $L
