#!/bin/bash
# Round 2, GPU call 14: the VAE route / seam tests with the TF32 switches at their defaults (earlier test files used to leave
# cuDNN's TF32 switch off for the rest of a one-process run, which kept the engine convolution out of these tests).
mkdir -p gpurun_out
L=gpurun_out/r2_call14.log
date > $L
timeout 280 python -m pytest tests/test_kernels_gpu.py tests/test_seams_gpu.py -q -m gpu -s -p no:cacheprovider --timeout 250 -k "vae or groupnorm_fp32 or f16_operands or pipeline or serving" >> $L 2>&1
echo "exit $?" >> $L
grep -n "passed\|failed\|FAILED\|VAE vs\|B1\|exit" $L | tail -n 12
