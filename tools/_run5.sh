mkdir -p gpurun_out
timeout 120 ./tools/dfma_bench > gpurun_out/r3_dfma_bench.json 2>&1; echo rc=$?
cat gpurun_out/r3_dfma_bench.json
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3_t3.log 2>&1; echo rc=$? >> gpurun_out/r3_t3.log
tail -4 gpurun_out/r3_t3.log
timeout 900 python bench.py > gpurun_out/r3_bench_n1_ws.json 2> gpurun_out/r3_bench_n1_ws.err; echo rc=$?
cat gpurun_out/r3_bench_n1_ws.json
