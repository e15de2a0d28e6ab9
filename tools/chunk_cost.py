"""What do the time-slice chunks of the overlapped sum cost the kernels by themselves?  BASELINE configs[3] share on one GPU
(32^3x64, 125 eigenvectors, ultra-local + 8 one-hop loops), one-rank communicator (no bytes move): one launch against the
chunk schedules of nchunks = 4, 8, 16."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mugiq_b200 import ops, synth
from mugiq_b200.params import parse_disp_entries, which_displace

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29577")
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
L = tuple(int(x) for x in os.environ.get("CHUNK_L", "32,32,32,64").split(","))
nev = int(os.environ.get("CHUNK_NEV", "125"))
_, ds, a, b = parse_disp_entries(synth.ONE_HOP_ENTRIES)
entries = [which_displace(s) + (x, y) for s, x, y in zip(ds, a, b)]
U = synth.random_gauge_slab_torch(L, list(range(L[3])), seed=13, device="cuda")
gd = torch.stack(list(U)).contiguous()
ev = synth.random_evecs_torch(L, nev, seed=100)
sig = synth.sigmas(nev)
plan = ops.LoopPlan(gd, entries, L)
pos = torch.zeros((plan.nLoop, 16, ev.shape[1]), dtype=torch.complex128, device="cuda")
prep = plan.prepare(list(ev), sig)
comm = ops.Comm(dist.group.WORLD)
out = {}


def once(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


# round-robin (the SM clock drifts under the power cap: interleave the variants), medians over 7 rounds after a warm-up round
variants = {"one_launch_ms": lambda: plan.accumulate(pos, prep, accumulate=False)}
for n in (2, 4, 8, 16):
    variants[f"nchunks_{n}_ms"] = (lambda n=n: plan.accumulate_allreduce(pos, prep, comm, accumulate=False, nchunks=n))
times = {k: [] for k in variants}
for r in range(8):
    for k, fn in variants.items():
        t = once(fn)
        if r:
            times[k].append(t)
import statistics
out = {k: round(statistics.median(v), 3) for k, v in times.items()}
out["min"] = {k: round(min(v), 3) for k, v in times.items()}
print(json.dumps(out))
comm.close()
dist.destroy_process_group()
