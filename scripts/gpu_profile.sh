#!/bin/bash
# Profiling call: determinism re-check, per-kernel microbench, ncu launch list of one denoise step, ncu --set full of
# the dominant kernels. Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -s --timeout 600 -p no:cacheprovider -k "loop or shared" > gpurun_out/engine_tests2.log 2>&1; echo "engine tests exit $?"; tail -n 6 gpurun_out/engine_tests2.log )
( timeout 900 python scripts/microbench.py > gpurun_out/microbench.log 2>&1; echo "microbench exit $?"; cat gpurun_out/microbench.log | tail -n 80 )
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2100 -c 2100 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launch list exit $?"; wc -l gpurun_out/launches.csv )
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_conv|attn_kernel" -s 8 -c 4 -o gpurun_out/prof_r1 \
    python scripts/prof_target.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"; tail -n 3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep )
