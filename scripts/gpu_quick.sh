#!/bin/bash
mkdir -p gpurun_out
date > gpurun_out/quick.log
( timeout 420 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm2 or 2cta or auto" --timeout 200 -p no:cacheprovider -x >> gpurun_out/quick.log 2>&1; echo "gemm2 tests exit $?" | tee -a gpurun_out/quick.log; date >> gpurun_out/quick.log; tail -n 6 gpurun_out/quick.log )
if grep -q "gemm2 tests exit 0" gpurun_out/quick.log; then
( MB_ONLY=gemm timeout 300 python scripts/microbench.py > gpurun_out/microbench5.log 2>&1; echo "microbench exit $?"; grep -E "gemm2|conv3x3_2cta" gpurun_out/microbench5.log | grep -vE "bn.: (128|192)" | tail -n 40 )
fi
