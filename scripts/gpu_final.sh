#!/bin/bash
# round-end validation: full GPU test suite, smoke(), default bench (value + e2e + cpu_baseline)
mkdir -p gpurun_out
( timeout 500 python -m pytest tests -q -m gpu --timeout 240 -p no:cacheprovider > gpurun_out/final_tests.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/final_tests.log; tail -n 6 gpurun_out/final_tests.log )
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/final_smoke.log )
( timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $?"; python -c "import json;d=json.load(open('gpurun_out/final_bench.json'));print('bench', d['value'], d['ms_per_step'], d['e2e'], d['cpu_baseline'], d['roofline']['frac'], d['roofline']['step']['frac'], d['clocks'], d['gpu_launches'])" )
