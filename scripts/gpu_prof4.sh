#!/bin/bash
mkdir -p gpurun_out
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm2_kernel|attn4_kernel|gn_apply|layernorm" -s 14 -c 7 -o gpurun_out/prof_r1_final \
    python scripts/prof_target3.py > gpurun_out/ncu_full4.log 2>&1; echo "ncu full exit $?"; tail -n 2 gpurun_out/ncu_full4.log; ls -la gpurun_out/prof_r1_final.ncu-rep )
( timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --profile-one-step --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench3.log 2>&1; echo "ncu launch list exit $?"; wc -l gpurun_out/launches_final.csv )
