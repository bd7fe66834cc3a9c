#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm2 or 2cta or auto" --timeout 120 -p no:cacheprovider > gpurun_out/kt_gemm2.log 2>&1; echo "gemm2 tests exit $?"; tail -n 30 gpurun_out/kt_gemm2.log )
if grep -q " passed" gpurun_out/kt_gemm2.log && ! grep -q "failed" gpurun_out/kt_gemm2.log; then
  export V2OK=1
else
  export B200VTON_GEMM2=0 MB_V2=0
fi
( MB_ONLY=gemm timeout 600 python scripts/microbench.py > gpurun_out/microbench2.log 2>&1; echo "microbench exit $?"; grep -E "gemm2|conv3x3_2cta|cublas|cudnn" gpurun_out/microbench2.log | tail -n 70 )
( timeout 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -s --timeout 600 -p no:cacheprovider > gpurun_out/engine_tests3.log 2>&1; echo "engine tests exit $?"; grep -E "golden:|eps eng|loop 3|shared|passed|failed" gpurun_out/engine_tests3.log )
( timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --profile-one-step --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launch list exit $?"; wc -l gpurun_out/launches.csv )
( timeout 900 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench exit $?"; tail -n 3 gpurun_out/bench2.err; cat gpurun_out/bench2.json )
