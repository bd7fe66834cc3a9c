#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm2 or 2cta or auto or attention" --timeout 300 -p no:cacheprovider -s > gpurun_out/kt3.log 2>&1; echo "kernel tests exit $?"; tail -n 25 gpurun_out/kt3.log )
( timeout 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -s --timeout 600 -p no:cacheprovider > gpurun_out/engine_tests5.log 2>&1; echo "engine tests exit $?"; grep -E "golden:|eps eng|loop 3|shared|hoisted|passed|failed|Error" gpurun_out/engine_tests5.log )
( timeout 600 python scripts/microbench.py > gpurun_out/microbench4.log 2>&1; echo "microbench exit $?"; grep -E "gemm2|conv3x3_2cta|attention|sdpa" gpurun_out/microbench4.log | grep -vE "bn.: 128" | tail -n 50 )
( timeout 900 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench exit $?"; tail -n 3 gpurun_out/bench4.err; cat gpurun_out/bench4.json )
( timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches2.csv \
    python bench.py --profile-one-step --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1; echo "ncu launch list exit $?"; wc -l gpurun_out/launches2.csv )
