#!/bin/bash
mkdir -p gpurun_out
( timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_r1_v5 -f \
    python scripts/prof_target4.py > gpurun_out/ncu_full5.log 2>&1; echo "ncu full exit $?"; tail -n 2 gpurun_out/ncu_full5.log; ls -la gpurun_out/prof_r1_v5.ncu-rep )
