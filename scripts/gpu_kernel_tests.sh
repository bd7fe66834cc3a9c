#!/bin/bash
# Runs the per-kernel GPU parity tests in separate processes (a trapped kernel poisons its CUDA context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for grp in gemm conv3x3 attention "groupnorm or layernorm" "layout or downsample or timestep or cfg"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$grp" --timeout 120 -p no:cacheprovider \
      > "gpurun_out/kt_${name}.log" 2>&1
  echo "== $grp: exit $?"; tail -n 25 "gpurun_out/kt_${name}.log"
done
