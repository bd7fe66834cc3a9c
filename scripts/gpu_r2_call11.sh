#!/bin/bash
# Round 2, GPU call 11: first point of divergence in the serving scenario (loop inputs, context K/V, garment K/V, per-step
# eps / latents), with the step graph + PDL, the graph without PDL, and eager launches.
mkdir -p gpurun_out
L=gpurun_out/r2_call11.log
date > $L
run() { echo "=== $*" >> $L; env "$@" timeout 300 python scripts/diag_serving_determinism.py $ARGS 2>&1 | tail -n 2 >> $L; }
ARGS="" run B200VTON_CLIP=1
ARGS="" run B200VTON_CLIP=1
ARGS="" run B200VTON_CLIP=1 B200VTON_PDL_GRAPH=0
ARGS="" run B200VTON_CLIP=1 B200VTON_PDL_GRAPH=0
ARGS="--eager" run B200VTON_CLIP=1
ARGS="--eager" run B200VTON_CLIP=1
ARGS="" run B200VTON_CLIP=0
ARGS="" run B200VTON_CLIP=1 B200VTON_VAE_NHWC=0
ARGS="" run B200VTON_CLIP=1 CUDA_LAUNCH_BLOCKING=1
cat $L
