#!/bin/bash
# Round 2, validation of the VAE fusions (residual add in the TF32 convolution's epilogue; one-pass split / softmax-split
# kernels + 3x-long contraction in the mid-block attention): the driver's sequence + a VAE A/B timing.
mkdir -p gpurun_out
L=gpurun_out/r2_final5.log
date > $L
step() { echo "=== $1" | tee -a $L; shift; ( "$@" ) >> $L 2>&1; echo "    exit $?" | tee -a $L; }
step "pytest tests -m gpu (one process, as the driver runs it)" timeout 1500 python -m pytest tests/ -x -q -m gpu -s --timeout 900 -p no:cacheprovider
step "smoke" timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== VAE A/B" | tee -a $L
timeout 300 python scripts/vae_timing.py > gpurun_out/r2_vae_fused_ab.json 2>> $L; echo "    exit $?" | tee -a $L
cat gpurun_out/r2_vae_fused_ab.json >> $L
echo "=== bench default (stage trace)" | tee -a $L
B200VTON_TRACE=1 timeout 600 python bench.py > gpurun_out/r2_bench_final5.json 2> gpurun_out/r2_bench_final5.err; echo "    exit $?" | tee -a $L
grep "b200vton trace" gpurun_out/r2_bench_final5.err | tail -n 2 >> $L
grep -n "passed\|failed\|smoke:\|exit\|VAE attention\|softmax_split\|_ms" $L | tail -n 30
