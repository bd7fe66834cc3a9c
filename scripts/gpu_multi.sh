#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline ${EXTRA} > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"; grep -vE "^W|warn|OMP" gpurun_out/bench_n$N.err | tail -n 8; cat gpurun_out/bench_n$N.json )
