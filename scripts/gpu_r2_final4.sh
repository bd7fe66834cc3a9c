#!/bin/bash
# Round 2, final validation after the CLIP towers / serving fix: the driver's sequence (whole GPU suite in one process,
# smoke(), default bench) + one ncu --set full capture of the new encoder-attention kernel.
mkdir -p gpurun_out
L=gpurun_out/r2_final4.log
date > $L
step() { echo "=== $1" | tee -a $L; shift; ( "$@" ) >> $L 2>&1; echo "    exit $?" | tee -a $L; }
step "pytest tests -m gpu (one process, as the driver runs it)" timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 900 -p no:cacheprovider
step "serving test x3 (was flaky)" bash -c 'for i in 1 2 3; do timeout 300 python -m pytest tests/test_seams_gpu.py -q -k serving -p no:cacheprovider 2>&1 | tail -n 1; done'
step "smoke" timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== bench default" | tee -a $L
timeout 600 python bench.py > gpurun_out/r2_bench_final4.json 2> gpurun_out/r2_bench_final4.err; echo "    exit $?" | tee -a $L
tail -n 3 gpurun_out/r2_bench_final4.err >> $L
step "ncu enc_attn" timeout 300 ncu --set full --clock-control none --import-source on -k regex:enc_attn --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2_enc_attn_vith python scripts/prof_target_clip.py
grep -n "passed\|failed\|smoke:\|exit" $L | tail -n 20
tail -n 12 $L
