#!/bin/bash
# Round 2, GPU call 10: where does the serving scenario lose bit-reproducibility (CLIP on the engine vs the module;
# with / without NaN-poisoned free memory)?
mkdir -p gpurun_out
L=gpurun_out/r2_call10.log
date > $L
for clip in 1 0; do for p in "" "--poison"; do
  echo "=== CLIP=$clip $p" >> $L
  B200VTON_CLIP=$clip timeout 300 python scripts/diag_serving_determinism.py $p 2>&1 | tail -n 3 >> $L
done; done
echo "=== serving test x2, CLIP=1" >> $L
for i in 1 2; do timeout 300 python -m pytest tests/test_seams_gpu.py -q -k serving -p no:cacheprovider 2>&1 | tail -n 2 >> $L; done
echo "=== serving test x2, CLIP=0" >> $L
for i in 1 2; do B200VTON_CLIP=0 timeout 300 python -m pytest tests/test_seams_gpu.py -q -k serving -p no:cacheprovider 2>&1 | tail -n 2 >> $L; done
cat $L
