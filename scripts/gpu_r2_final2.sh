mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/ -x -q -m gpu --timeout 900 -p no:cacheprovider 2>&1 | tail -n 8 ) > gpurun_out/r2_final2.log 2>&1
( timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -s -k "3xtf32" -p no:cacheprovider 2>&1 | grep "VAE attention" ) >> gpurun_out/r2_final2.log 2>&1
B200VTON_TRACE=1 timeout 500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_final2_bench.json 2> gpurun_out/r2_final2_bench.err
grep "b200vton trace" gpurun_out/r2_final2_bench.err | tail -n 1 >> gpurun_out/r2_final2.log
cat gpurun_out/r2_final2.log
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_final2_bench.json') if l.startswith('{')][0]); print('BENCH', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_call'])"
