"""Where the host spends its time inside one Loop_Mugiq.computeCoarseLoop step of BASELINE configs[1] (the GPU idles until
the first kernel is launched): perf_counter stamps around the phases of the step, median over 20 steps.  One JSON object."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mugiq_b200 import ops, synth
from mugiq_b200.loop import Loop_Mugiq, Eigsolve
from mugiq_b200.params import MugiqLoopParam, momenta_up_to

L, nev = (16, 16, 16, 32), 200
U = synth.random_gauge(L, seed=11)
ev = synth.random_evecs_torch(L, nev, seed=100)
sig = synth.sigmas(nev)
prm = MugiqLoopParam(gauge=[U[mu] for mu in range(4)])
prm.set_displacements(synth.ONE_HOP_ENTRIES)
prm.set_momenta(momenta_up_to(1))
loop = Loop_Mugiq(prm, Eigsolve(list(ev), sig, L), evec_batch=200, copy_pos_to_host=False)
stamps = []
orig = {}


def wrap(obj, name, label):
    f = getattr(obj, name)
    orig[label] = f

    def g(*a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        stamps[-1].append((label, t0, time.perf_counter()))
        return r
    setattr(obj, name, g)


wrap(loop, "_loop_plan", "loop_plan")
wrap(loop, "_prepared_batch", "prepared_batch")
wrap(loop, "performMomentumProjection", "projection+D2H")
for _ in range(3):
    stamps.append([])
    loop.MomProjDone = False
    loop.computeCoarseLoop()
plan = loop._plan
wrap(plan, "accumulate", "plan.accumulate (fused launch)")
wrap(plan, "finalize", "plan.finalize")
rows = []
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for it in range(20):
    torch.cuda.synchronize()
    stamps.append([])
    t0 = time.perf_counter()
    e[0].record()
    loop.MomProjDone = False
    loop.computeCoarseLoop()
    t1 = time.perf_counter()
    e[1].record()
    torch.cuda.synchronize()
    d = {lab: (b - a) * 1e6 for lab, a, b in stamps[-1]}
    d["launch_of_fused_kernel_at_us"] = [a for lab, a, b in stamps[-1] if lab.startswith("plan.accumulate")][0] * 1e6 - t0 * 1e6
    d["fused_launch_returns_at_us"] = [b for lab, a, b in stamps[-1] if lab.startswith("plan.accumulate")][0] * 1e6 - t0 * 1e6
    d["step_host_us"] = (t1 - t0) * 1e6
    d["step_gpu_events_us"] = e[0].elapsed_time(e[1]) * 1e3
    rows.append(d)
out = {k: round(float(np.median([r[k] for r in rows])), 1) for k in rows[0]}
print(json.dumps(out))
