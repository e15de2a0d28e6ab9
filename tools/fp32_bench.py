"""FP32 instantiation of the path at BASELINE configs[1] (16^3x32, 200 eigenvectors, ultra-local + 8 one-hop loops, 7 momenta):
step time and per-kernel times from the library's event timers.  FP32 is not the headline precision (the metric is quoted
in FP64); this is the one measurement of it.  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mugiq_b200 import ops, synth
from mugiq_b200.loop import Loop_Mugiq, Eigsolve
from mugiq_b200.params import MugiqLoopParam, momenta_up_to

L, nev = (16, 16, 16, 32), 200
U = synth.random_gauge(L, seed=11).astype(np.complex64)
prm = MugiqLoopParam(gauge=[U[mu] for mu in range(4)])
prm.set_displacements(synth.ONE_HOP_ENTRIES)
prm.set_momenta(momenta_up_to(1))
ev = synth.random_evecs_torch(L, nev, seed=100, dtype=torch.complex64)
sig = synth.sigmas(nev)
loop = Loop_Mugiq(prm, Eigsolve(list(ev), sig, L), evec_batch=200, copy_pos_to_host=False)


def step():
    loop.MomProjDone = False
    loop.computeCoarseLoop()


for _ in range(3):
    step()
torch.cuda.synchronize()
ops.prof_reset()
ops.prof_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 10
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
ops.prof_enable(False)
ms = e0.elapsed_time(e1) / steps
rep = ops.prof_report()
V4 = int(np.prod(L))
s_tot = float((1.0 / sig).sum())
print(json.dumps({"precision": "f32", "workload": "16x16x16x32_nev200_ulocal+1hop8", "ms_per_step": ms,
                  "value": nev * V4 * 9 / (ms * 1e-3), "kernels_ms_per_step": {k: v["ms"] / steps for k, v in rep.items()},
                  "loop_fused_GBps": rep["loop_fused"]["alg_bytes"] / (rep["loop_fused"]["ms"] * 1e-3) / 1e9,
                  "loop_fused_TFLOPs_fp32": rep["loop_fused"]["alg_flops"] / (rep["loop_fused"]["ms"] * 1e-3) / 1e12,
                  "checksum_rel_err": abs(complex(loop.dataPos_d[0, 0].sum().item()) - s_tot) / s_tot}))
