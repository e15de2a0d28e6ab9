mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --tsplit --steps 5 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/r3_tsplit_small_n2_ramp.json 2> gpurun_out/r3_tsplit_small_n2_ramp.err; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r3_tsplit_small_n2_ramp.json')); print(d['ms_per_step'], d['roofline']['kernels'], d['verified'])"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r3_bench_n2_b.json 2> gpurun_out/r3_bench_n2_b.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/r3_bench_n2_b.json"))
print(d["ms_per_step"], d["value"])
print("config4", d["config4"]["ms_per_step"])
print("tsplit", d["tsplit"])
PY
timeout 600 python -m pytest tests/test_tsplit.py tests/test_host_mirror.py -m gpu -x -q 2>&1 | tail -2
