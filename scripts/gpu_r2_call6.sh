#!/bin/bash
# Round 2, GPU call 6: three-part pipeline golden test, garment-chunk sweep (timesteps batched per hoisted garment pass).
mkdir -p gpurun_out
L=gpurun_out/r2_call6.log
date > $L
step() { echo "=== $1" | tee -a $L; shift; ( "$@" ) >> $L 2>&1; echo "    exit $?" | tee -a $L; }
step "pipeline golden + serving" timeout 600 python -m pytest tests/test_seams_gpu.py -q -m gpu -s --timeout 500 -p no:cacheprovider -k "pipeline or serving"
for c in 8 15 30; do
  echo "=== bench loop only, garment chunk = $c" | tee -a $L
  B200VTON_GARMENT_CHUNK=$c timeout 400 python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_bench_chunk_$c.json 2> gpurun_out/r2_bench_chunk_$c.err; echo "    exit $?" | tee -a $L
  tail -n 2 gpurun_out/r2_bench_chunk_$c.err >> $L
done
grep -h '"metric"' gpurun_out/r2_bench_chunk_*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('BENCH value', round(d['value'], 4), 'ms/loop', round(d['ms_per_step'],1), 'eager launches', d['launches_eager_per_bench_step'])" | tee -a $L
tail -n 60 $L
