"""One launch of the fused kernel per variant at BASELINE configs[1] (16^3x32, 200 eigenvectors, ultra-local + 8 one-hop
loops): eigenvectors in the canonical site-major order, then in QUDA FLOAT2 order (staged as TMA tensor boxes).
Target of `ncu --set full -k regex:loop_fused` (tools/ncu_digest.py reads the report); prints the CUDA-event times."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mugiq_b200 import ops, synth
from mugiq_b200.params import parse_disp_entries, which_displace

L = tuple(int(x) for x in os.environ.get("NCU_FUSED_L", "16,16,16,32").split(","))
nev = int(os.environ.get("NCU_FUSED_NEV", "200"))
text = os.environ.get("NCU_FUSED_ENTRIES", synth.ONE_HOP_ENTRIES)
_, ds, a, b = parse_disp_entries(text)
entries = [which_displace(s) + (x, y) for s, x, y in zip(ds, a, b)]
U = synth.random_gauge(L, seed=11)
gd = ops.gauge_upload(U, L)
ev = synth.random_evecs_torch(L, nev, seed=100)
sig = synth.sigmas(nev)
out = {}
for name, order in (("site_major", 0), ("quda_float2", 2)):
    fields = ev if order == 0 else torch.stack([ops.export_spinor(ev[i], 2, L) for i in range(nev)])
    plan = ops.LoopPlan(gd, entries, L)
    plan.set_evec_order(order)
    pos = torch.zeros((plan.nLoop, 16, ev.shape[1]), dtype=torch.complex128, device="cuda")
    prep = plan.prepare(list(fields), sig)
    times = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.accumulate(pos, prep, accumulate=False)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    out[name] = {"ms": min(times), "checksum_rel_err": abs(complex(pos[0, 0].sum().item()) - float((1.0 / sig).sum())) / float((1.0 / sig).sum())}
    plan.close()
    del fields, pos
print(json.dumps(out))
