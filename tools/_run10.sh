mkdir -p gpurun_out
MUGIQ_B200_TSPLIT_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --tsplit --steps 5 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/r3_tsplit_small_n2.json 2> gpurun_out/r3_tsplit_small_n2.err; echo rc=$?
grep "T-split phases" gpurun_out/r3_tsplit_small_n2.err | head -4
python -c "
import json; d=json.load(open('gpurun_out/r3_tsplit_small_n2.json')); print(d['ms_per_step'], d['roofline']['kernels'])"
MUGIQ_B200_TSPLIT_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --tsplit --tsplit-batch 200 --steps 5 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/r3_tsplit_small_n2_b200.json 2> gpurun_out/r3_tsplit_small_n2_b200.err; echo rc=$?
grep "T-split phases" gpurun_out/r3_tsplit_small_n2_b200.err | head -4
python -c "
import json; d=json.load(open('gpurun_out/r3_tsplit_small_n2_b200.json')); print(d['ms_per_step'], d['roofline']['kernels'])"
