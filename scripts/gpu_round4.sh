#!/bin/bash
mkdir -p gpurun_out
( timeout 500 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 200 -p no:cacheprovider -x > gpurun_out/r4_kernels.log 2>&1; echo "exit $?" >> gpurun_out/r4_kernels.log; tail -n 5 gpurun_out/r4_kernels.log )
( timeout 400 python -m pytest tests/test_engine_gpu.py -q -m gpu --timeout 200 -p no:cacheprovider -x > gpurun_out/r4_engine.log 2>&1; echo "exit $?" >> gpurun_out/r4_engine.log; tail -n 5 gpurun_out/r4_engine.log )
( B200VTON_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err; echo "bench exit $?"; grep "b200vton trace" gpurun_out/r4_bench.err | tail -n 1; python -c "import json;d=json.load(open('gpurun_out/r4_bench.json'));print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_call'], d['clocks'])" )
