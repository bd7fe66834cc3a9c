#!/bin/bash
mkdir -p gpurun_out
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn3_kernel|gemm2_kernel" -s 3 -c 3 -o gpurun_out/prof_r1b \
    python scripts/prof_target2.py > gpurun_out/ncu_full2.log 2>&1; echo "ncu full exit $?"; tail -n 3 gpurun_out/ncu_full2.log; ls -la gpurun_out/prof_r1b.ncu-rep )
