#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 200 -p no:cacheprovider -x > gpurun_out/pdl_kernels.log 2>&1; echo "exit $?" >> gpurun_out/pdl_kernels.log; tail -n 4 gpurun_out/pdl_kernels.log )
( timeout 400 python -m pytest tests/test_engine_gpu.py -q -m gpu --timeout 200 -p no:cacheprovider -x > gpurun_out/pdl_engine.log 2>&1; echo "exit $?" >> gpurun_out/pdl_engine.log; tail -n 4 gpurun_out/pdl_engine.log )
( timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/pdl_bench_on.json 2> gpurun_out/pdl_bench_on.err; python -c "import json;d=json.load(open('gpurun_out/pdl_bench_on.json'));print('pdl on', d['value'], d['ms_per_step'])" )
( B200VTON_PDL=0 timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/pdl_bench_off.json 2> gpurun_out/pdl_bench_off.err; python -c "import json;d=json.load(open('gpurun_out/pdl_bench_off.json'));print('pdl off', d['value'], d['ms_per_step'])" )
