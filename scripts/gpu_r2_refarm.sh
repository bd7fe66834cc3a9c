mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_reference_arm.json 2> gpurun_out/r2_reference_arm.err
echo "exit $? after $(( $(date +%s) - S )) s"
grep -c "full" gpurun_out/r2_reference_arm.err; tail -n 4 gpurun_out/r2_reference_arm.err; cut -c1-900 gpurun_out/r2_reference_arm.json
