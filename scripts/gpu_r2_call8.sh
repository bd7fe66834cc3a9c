#!/bin/bash
# Round 2, GPU call 8: compute-sanitizer memcheck with the caching allocator OFF (every tensor its own cudaMalloc, so an
# out-of-bounds access of a few bytes past ANY tensor is caught — the time-embedding over-read of round 1 hid inside the
# allocator's blocks for a whole round), over the kernel, engine and seam tests; racecheck on the GroupNorm barrier kernel.
mkdir -p gpurun_out
L=gpurun_out/r2_call8.log
date > $L
step() { echo "=== $1" | tee -a $L; shift; ( "$@" ) >> $L 2>&1; echo "    exit $?" | tee -a $L; }
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
step "memcheck kernels (no caching allocator)" timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider --timeout 1400 -x
step "memcheck engine (no caching allocator)" timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_engine_gpu.py -q -m gpu -p no:cacheprovider --timeout 1400 -x
step "memcheck seams (no caching allocator)" timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_seams_gpu.py -q -m gpu -p no:cacheprovider --timeout 1400 -x
unset PYTORCH_NO_CUDA_MEMORY_CACHING
step "racecheck GroupNorm" timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider --timeout 500 -k "groupnorm and not fp32"
grep -n "ERROR SUMMARY\|passed\|failed\|Invalid\|Race" $L | head -40
tail -n 40 $L
