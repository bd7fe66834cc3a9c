#!/bin/bash
# Round 2, GPU call 5: the driver's own sequence — whole GPU suite in one process, smoke(), default bench.
mkdir -p gpurun_out
L=gpurun_out/r2_call5.log
date > $L
rm -f gpurun_out/fullsize_parity.jsonl
step() { echo "=== $1" | tee -a $L; shift; ( "$@" ) >> $L 2>&1; echo "    exit $?" | tee -a $L; }
step "pytest tests -m gpu (one process, as the driver runs it)" timeout 1500 python -m pytest tests/ -q -m gpu -s --timeout 900 -p no:cacheprovider
step "smoke" timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== bench default" | tee -a $L
timeout 600 python bench.py > gpurun_out/r2_bench_call5.json 2> gpurun_out/r2_bench_call5.err; echo "    exit $?" | tee -a $L
tail -n 4 gpurun_out/r2_bench_call5.err >> $L
grep -n "passed\|failed\|PARITY\|B1 __call__\|smoke:" $L | tail -n 20
tail -n 30 $L
