#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"; tail -n 12 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json )
