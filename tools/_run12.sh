mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3_t5.log 2>&1; echo rc=$? >> gpurun_out/r3_t5.log
tail -3 gpurun_out/r3_t5.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3_bench_reference_arm.json 2> gpurun_out/r3_bench_reference_arm.err; echo rc=$?
head -c 600 gpurun_out/r3_bench_reference_arm.json; echo
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra > gpurun_out/r3_bench_short.json 2> gpurun_out/r3_bench_short.err; echo rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3_ncu_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra > gpurun_out/r3_ncu_launches.log 2>&1; echo rc=$?
timeout 600 ncu --set full --import-source on --clock-control none -k regex:loop_fused -c 1 -o gpurun_out/r3_prof_final --force-overwrite python tools/ncu_fused.py > gpurun_out/r3_ncu_final.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none -k regex:"momproj_pos|minus_from_plus" -c 2 -o gpurun_out/r3_prof_aux --force-overwrite python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extra --no-verify > gpurun_out/r3_ncu_aux.log 2>&1; echo rc=$?
