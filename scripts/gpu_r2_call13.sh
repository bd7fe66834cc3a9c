#!/bin/bash
# Round 2, GPU call 13: ncu --set full of the VAE attention's one-pass kernels (achieved DRAM throughput of HBM-bound kernels).
mkdir -p gpurun_out
L=gpurun_out/r2_call13.log
date > $L
timeout 400 ncu --set full --clock-control none --import-source on -k regex:split_tf32 --launch-skip 6 --launch-count 4 -f -o gpurun_out/r2_vae_split_kernels python scripts/prof_target_vae.py >> $L 2>&1
echo "exit $?" >> $L
tail -n 5 $L
