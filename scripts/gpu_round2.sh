#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm2 or 2cta or auto" --timeout 120 -p no:cacheprovider > gpurun_out/kt_gemm2.log 2>&1; echo "gemm2 tests exit $?"; tail -n 12 gpurun_out/kt_gemm2.log )
( timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" --timeout 120 -p no:cacheprovider > gpurun_out/kt_attn.log 2>&1; echo "attention tests exit $?"; tail -n 30 gpurun_out/kt_attn.log )
grep -q "failed" gpurun_out/kt_gemm2.log && export B200VTON_GEMM2=0 MB_V2=0
grep -q "failed" gpurun_out/kt_attn.log && export B200VTON_ATTN2=0
( timeout 600 python scripts/microbench.py > gpurun_out/microbench3.log 2>&1; echo "microbench exit $?"; grep -E "gemm2|conv3x3_2cta|attention|sdpa" gpurun_out/microbench3.log | tail -n 70 )
( timeout 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -s --timeout 600 -p no:cacheprovider > gpurun_out/engine_tests4.log 2>&1; echo "engine tests exit $?"; grep -E "golden:|eps eng|loop 3|shared|passed|failed|Error" gpurun_out/engine_tests4.log )
( timeout 900 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench exit $?"; tail -n 3 gpurun_out/bench3.err; cat gpurun_out/bench3.json )
