#!/bin/bash
# Round 2, GPU call 2: whole GPU suite (new seam / full-size tests included), PDL validation (stress + synccheck),
# bench at BASELINE configs 2, 3 and 4.
mkdir -p gpurun_out
L=gpurun_out/r2_call2.log
date > $L
rm -f gpurun_out/fullsize_parity.jsonl
step() { echo "=== $1" | tee -a $L; shift; ( "$@" ) >> $L 2>&1; echo "    exit $?" | tee -a $L; }
step "GPU test suite" timeout 1500 python -m pytest tests -q -m gpu -s --timeout 600 -p no:cacheprovider
for i in 1 2 3 4; do
  step "PDL stress run $i (bench with e2e, programmatic dependent launch ON)" env B200VTON_PDL=1 B200VTON_E2E_TIMEOUT=150 timeout 300 python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-eager-baseline
done
step "synccheck with PDL on (tiny loop + graph)" env B200VTON_PDL=1 timeout 420 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_engine_gpu.py -q -m gpu -k "tiny_loop_and_graph" -p no:cacheprovider --timeout 400
step "memcheck (tiny UNets vs golden)" timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_engine_gpu.py -q -m gpu -k "reference_golden" -p no:cacheprovider --timeout 400
step "BN sweep" timeout 300 python scripts/mb_bn.py
cp gpurun_out/r2_call2.log gpurun_out/r2_call2_partial.log
echo "=== bench config 2 (default invocation)" | tee -a $L
timeout 600 python bench.py > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err; echo "    exit $?" | tee -a $L
echo "=== bench config 3" | tee -a $L
timeout 600 python bench.py --config 3 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench_cfg3.json 2> gpurun_out/r2_bench_cfg3.err; echo "    exit $?" | tee -a $L
echo "=== bench config 4" | tee -a $L
timeout 900 python bench.py --config 4 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err; echo "    exit $?" | tee -a $L
tail -n 5 gpurun_out/r2_bench_cfg*.err >> $L
grep -h '"metric"' $L gpurun_out/r2_bench_cfg*.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('BENCH', d['config'].get('height'), d['config'].get('batch_per_loop'), 'value', round(d['value'], 3), 'ms', round(d['ms_per_step'],1), 'e2e', d['e2e'] and d['e2e'].get('value'), 'eager', d.get('eager_gpu_baseline') and d['eager_gpu_baseline'].get('value'), 'step_frac', round(d['roofline']['step']['frac'],3))" | tee -a $L
tail -n 80 $L
