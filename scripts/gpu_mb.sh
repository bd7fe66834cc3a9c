#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm2 or 2cta or auto" --timeout 300 -p no:cacheprovider > gpurun_out/kt_mb.log 2>&1; echo "gemm2 tests exit $?"; tail -n 4 gpurun_out/kt_mb.log )
( MB_ONLY=gemm timeout 600 python scripts/microbench.py > gpurun_out/microbench5.log 2>&1; echo "microbench exit $?"; grep -E "gemm2|conv3x3_2cta" gpurun_out/microbench5.log | grep -vE "bn.: (128|192)" | tail -n 50 )
( timeout 900 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo "bench exit $?"; tail -n 3 gpurun_out/bench5.err; python -c "
import json; d=json.load(open('gpurun_out/bench5.json')); print({k: d[k] for k in ('value','ms_per_step','roofline','clocks')})" )
