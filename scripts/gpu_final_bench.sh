#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
( timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $? after $(( $(date +%s) - T0 )) s"; python -c "import json;d=json.load(open('gpurun_out/final_bench.json'));print('bench', d['value'], d['ms_per_step'], d['e2e'], d['cpu_baseline'], d['roofline']['frac'], d['roofline']['step']['frac'], d['clocks'], d['gpu_launches'])" )
