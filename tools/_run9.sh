mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r3_n2_gpus.txt
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r3_pytest_gpu_2gpus.log 2>&1; echo rc=$? >> gpurun_out/r3_pytest_gpu_2gpus.log
tail -8 gpurun_out/r3_pytest_gpu_2gpus.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r3_bench_n2.json 2> gpurun_out/r3_bench_n2.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/r3_bench_n2.json"))
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"].get("frac_of_pcie_ceiling"))
print("config4", {k:d["config4"].get(k) for k in ("ms_per_step","value","allreduce_ms","efficiency_vs_n1")}, d["config4"].keys())
print("tsplit", d["tsplit"])
PY
