#!/bin/bash
mkdir -p gpurun_out
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm2_kernel" -s 2 -c 2 -o gpurun_out/prof_r1c \
    python scripts/prof_target2.py > gpurun_out/ncu_full3.log 2>&1; echo "ncu full exit $?"; tail -n 3 gpurun_out/ncu_full3.log; ls -la gpurun_out/prof_r1c.ncu-rep )
