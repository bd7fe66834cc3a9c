#!/bin/bash
# Round 2 final validation: the driver's sequence (GPU suite in one process, smoke(), default bench) plus the other
# single-GPU configs and the multi-batch path of config 5 on one GPU.
mkdir -p gpurun_out
L=gpurun_out/r2_final.log
date > $L
rm -f gpurun_out/fullsize_parity.jsonl
step() { echo "=== $1" | tee -a $L; shift; ( "$@" ) >> $L 2>&1; echo "    exit $?" | tee -a $L; }
step "pytest tests -m gpu (one process, as the driver runs it)" timeout 1500 python -m pytest tests/ -x -q -m gpu -s --timeout 900 -p no:cacheprovider
step "smoke" timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== bench default (config 2)" | tee -a $L
B200VTON_TRACE=1 timeout 600 python bench.py > gpurun_out/r2_final_bench_cfg2.json 2> gpurun_out/r2_final_bench_cfg2.err; echo "    exit $?" | tee -a $L
grep "b200vton trace" gpurun_out/r2_final_bench_cfg2.err | tail -n 1 >> $L
echo "=== bench config 3" | tee -a $L
timeout 600 python bench.py --config 3 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2_final_bench_cfg3.json 2> gpurun_out/r2_final_bench_cfg3.err; echo "    exit $?" | tee -a $L
echo "=== bench config 4" | tee -a $L
timeout 900 python bench.py --config 4 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2_final_bench_cfg4.json 2> gpurun_out/r2_final_bench_cfg4.err; echo "    exit $?" | tee -a $L
echo "=== bench config 5 shape on ONE GPU, 16 requests = two batches of 8 per bench step" | tee -a $L
timeout 900 python bench.py --config 5 --requests 16 --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_final_bench_cfg5_n1.json 2> gpurun_out/r2_final_bench_cfg5_n1.err; echo "    exit $?" | tee -a $L
tail -n 3 gpurun_out/r2_final_bench_cfg*.err >> $L
grep -h '"metric"' gpurun_out/r2_final_bench_cfg*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); c = d['config']
    print('BENCH', c['height'], c['width'], 'B', c['batch_per_loop'], 'req', c['requests_total'], 'value', round(d['value'], 3), 'ms', round(d['ms_per_step'], 1), 'e2e', d['e2e'] and round(d['e2e']['value'], 3), 'eager', d.get('eager_gpu_baseline') and round(d['eager_gpu_baseline']['value'], 3), 'step_frac', round(d['roofline']['step']['frac'], 3), 'dom', round(d['roofline']['achieved']), 'window', c['garment_kv_window_steps'])" | tee -a $L
grep -n "passed\|failed\|smoke:" $L | tail -n 6
tail -n 25 $L
