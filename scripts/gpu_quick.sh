#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" --timeout 120 -p no:cacheprovider -x > gpurun_out/quick.log 2>&1; echo "exit $?" >> gpurun_out/quick.log; tail -n 6 gpurun_out/quick.log )
( timeout 400 python -m pytest tests/test_engine_gpu.py -q -m gpu --timeout 200 -p no:cacheprovider -x > gpurun_out/quick_engine.log 2>&1; echo "exit $?" >> gpurun_out/quick_engine.log; tail -n 6 gpurun_out/quick_engine.log )
( timeout 400 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; tail -c 1500 gpurun_out/quick_bench.json )
