#!/bin/bash
# Round 2: multi-GPU runs under torch.distributed.run (one rank per GPU). NGPU=8 CONFIG=5 -> BASELINE config 5
# (64 requests sharded 8-way, weights broadcast over NCCL on the flat arenas); CONFIG=2 -> N replicas of config 2.
N=${NGPU:-8}
C=${CONFIG:-5}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --config $C --steps ${STEPS:-2} --warmup ${WARMUP:-1} --no-cpu-baseline --no-eager-baseline \
    > gpurun_out/r2_bench_n${N}_cfg${C}.json 2> gpurun_out/r2_bench_n${N}_cfg${C}.err; echo "bench N=$N config $C exit $?" )
tail -n 12 gpurun_out/r2_bench_n${N}_cfg${C}.err
cat gpurun_out/r2_bench_n${N}_cfg${C}.json | tail -n 2 | cut -c1-1500
