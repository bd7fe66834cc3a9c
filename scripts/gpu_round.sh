#!/bin/bash
# One GPU call: engine parity tests, smoke, bench line (with e2e + cpu baseline).
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu -s --timeout 600 -p no:cacheprovider > gpurun_out/engine_tests.log 2>&1; echo "engine tests exit $?"; tail -n 30 gpurun_out/engine_tests.log )
( timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 5 gpurun_out/smoke.log )
( timeout 1200 python bench.py --steps ${BENCH_STEPS:-2} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -n 25 gpurun_out/bench.err; cat gpurun_out/bench.json )
