mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r3_bench_n8.json 2> gpurun_out/r3_bench_n8.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/r3_bench_n8.json"))
print(d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"].get("frac_of_pcie_ceiling"), d["e2e"].get("pcie_ceiling_gbs"))
c=d["config4"]; print("config4", c["ms_per_step"], c["value"], c.get("pos_allreduce_overlapped"), c.get("pos_allreduce_after_kernels"), c.get("without_pos_allreduce"))
print("tsplit", d["tsplit"])
print(d["clocks"])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --tsplit --steps 5 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/r3_bench_n8_tsplit_small.json 2> gpurun_out/r3_bench_n8_tsplit_small.err; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r3_bench_n8_tsplit_small.json')); print('small slabs T split', d['ms_per_step'], d['value'], d['roofline']['kernels'])"
timeout 1200 python -m pytest tests -m gpu -q -rs > gpurun_out/r3_pytest_gpu_8gpus.log 2>&1; echo rc=$? >> gpurun_out/r3_pytest_gpu_8gpus.log
tail -4 gpurun_out/r3_pytest_gpu_8gpus.log
