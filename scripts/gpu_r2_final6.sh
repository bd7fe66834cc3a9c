#!/bin/bash
# Round 2, validation of the VAE's fp16 GroupNorm -> convolution hand-off (b200vton_conv3x3_nhwc_f16in_f32): whole GPU suite,
# smoke(), VAE A/B, default bench with the stage trace.
mkdir -p gpurun_out
L=gpurun_out/r2_final6.log
date > $L
step() { echo "=== $1" | tee -a $L; shift; ( "$@" ) >> $L 2>&1; echo "    exit $?" | tee -a $L; }
step "pytest tests -m gpu (one process)" timeout 1500 python -m pytest tests/ -q -m gpu -s --timeout 900 -p no:cacheprovider
step "smoke" timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== VAE A/B" | tee -a $L
timeout 300 python scripts/vae_timing.py > gpurun_out/r2_vae_f16_ab.json 2>> $L; echo "    exit $?" | tee -a $L
cat gpurun_out/r2_vae_f16_ab.json >> $L
echo "=== bench default (stage trace)" | tee -a $L
B200VTON_TRACE=1 timeout 600 python bench.py > gpurun_out/r2_bench_final6.json 2> gpurun_out/r2_bench_final6.err; echo "    exit $?" | tee -a $L
grep "b200vton trace" gpurun_out/r2_bench_final6.err | tail -n 2 >> $L
grep -n "passed\|failed\|FAILED\|smoke:\|exit\|fp16-operand\|VAE vs\|_ms" $L | tail -n 40
