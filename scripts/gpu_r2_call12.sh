#!/bin/bash
# Round 2, GPU call 12: compute-sanitizer on the CLIP kernels (memcheck without the caching allocator, synccheck), ncu launch
# lists of the VAE (which kernels its time goes to) and of one ViT-H tower pass.
mkdir -p gpurun_out
L=gpurun_out/r2_call12.log
date > $L
step() { echo "=== $1" | tee -a $L; shift; ( "$@" ) >> $L 2>&1; echo "    exit $?" | tee -a $L; }
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
step "memcheck CLIP kernels + ViT-H tower (no caching allocator)" timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_clip_gpu.py -q -m gpu -p no:cacheprovider --timeout 800 -x -k "not text_towers"
unset PYTORCH_NO_CUDA_MEMORY_CACHING
step "synccheck CLIP kernels" timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_clip_gpu.py -q -m gpu -p no:cacheprovider --timeout 500 -x -k "encoder_attention or quick_gelu or patchify"
step "ncu launch list VAE" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv --log-file gpurun_out/r2_vae_launches.csv python scripts/prof_target_vae.py
step "ncu launch list ViT-H tower" timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vton --launch-skip 700 -c 300 --csv --log-file gpurun_out/r2_clip_launches.csv python scripts/clip_timing.py
grep -n "passed\|failed\|ERROR SUMMARY\|exit" $L | tail -n 20
