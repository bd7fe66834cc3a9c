#!/bin/bash
# Round 2, GPU call 9: CLIP towers on the engine's kernels (attn_enc.cu, clip.py) — kernel + tower parity tests, the
# pipeline seam tests (their tiny CLIP now runs on the engine), tower timing, e2e bench with the stage trace.
mkdir -p gpurun_out
L=gpurun_out/r2_call9.log
date > $L
step() { echo "=== $1" | tee -a $L; shift; ( "$@" ) >> $L 2>&1; echo "    exit $?" | tee -a $L; }
step "pytest test_clip_gpu" timeout 900 python -m pytest tests/test_clip_gpu.py -q -s --timeout 600 -p no:cacheprovider
step "pytest test_seams_gpu" timeout 900 python -m pytest tests/test_seams_gpu.py -q -s --timeout 600 -p no:cacheprovider
echo "=== clip timing" | tee -a $L
timeout 600 python scripts/clip_timing.py > gpurun_out/r2_clip_timing.json 2>> $L; echo "    exit $?" | tee -a $L
cat gpurun_out/r2_clip_timing.json >> $L
echo "=== bench default with stage trace" | tee -a $L
B200VTON_TRACE=1 timeout 600 python bench.py > gpurun_out/r2_bench_call9.json 2> gpurun_out/r2_bench_call9.err; echo "    exit $?" | tee -a $L
grep -n "clip\|vae\|denois\|e2e" gpurun_out/r2_bench_call9.err | tail -n 20 >> $L
grep -n "passed\|failed\|Error\|error\|ViT\|text\|encoder_attention" $L | tail -n 60
tail -n 25 $L
