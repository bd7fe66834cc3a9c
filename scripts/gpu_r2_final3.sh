mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/ -q -m gpu --timeout 900 -p no:cacheprovider 2>&1 | tail -n 12 ) > gpurun_out/r2_final3.log 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 ) >> gpurun_out/r2_final3.log 2>&1
cat gpurun_out/r2_final3.log
