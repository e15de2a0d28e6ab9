"""Per-CTA timeline of one launch of the fused kernel (mugiq_b200_prof_fused_trace): where a CTA's time goes (stage map,
prologue, eigenvector loop, epilogue), how long an SM sits between two CTAs, and the tail of the launch.  Prints one JSON
object.  Environment: FUSED_TRACE_L (lattice, default BASELINE configs[1]), FUSED_TRACE_NEV, FUSED_TRACE_ENTRIES."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mugiq_b200 import ops, synth
from mugiq_b200.params import parse_disp_entries, which_displace

L = tuple(int(x) for x in os.environ.get("FUSED_TRACE_L", "16,16,16,32").split(","))
nev = int(os.environ.get("FUSED_TRACE_NEV", "200"))
text = os.environ.get("FUSED_TRACE_ENTRIES", synth.ONE_HOP_ENTRIES)
_, ds, a, b = parse_disp_entries(text)
entries = [which_displace(s) + (x, y) for s, x, y in zip(ds, a, b)]
U = synth.random_gauge(L, seed=11)
gd = ops.gauge_upload(U, L)
ev = synth.random_evecs_torch(L, nev, seed=100)
sig = synth.sigmas(nev)
plan = ops.LoopPlan(gd, entries, L)
pos = torch.zeros((plan.nLoop, 16, ev.shape[1]), dtype=torch.complex128, device="cuda")
prep = plan.prepare(list(ev), sig)
for _ in range(2):
    plan.accumulate(pos, prep, accumulate=False)
torch.cuda.synchronize()
ncta_max = 1 << 16
buf = torch.zeros(16 * ncta_max, dtype=torch.int64, device="cuda")
ops.prof_fused_trace(buf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
plan.accumulate(pos, prep, accumulate=False)
e1.record()
torch.cuda.synchronize()
ops.prof_fused_trace(None)
t = buf.cpu().numpy().reshape(-1, 16)
t = t[t[:, 1] > 0]
# several fused launches (groups) share the buffer: the last launch wins per CTA index; keep CTAs of the longest launch
sm, t0, tmap, tloop0, tloop1, tend = (t[:, k] for k in range(6))
start = t0.min()
out = {"lattice": list(L), "nev": nev, "ctas_traced": int(len(t)), "ms_launch_events": e0.elapsed_time(e1),
       "span_ms": float(tend.max() - start) / 1e6}


def stats(x):
    x = np.asarray(x, dtype=np.float64) / 1e3
    return {"mean_us": round(float(x.mean()), 2), "p50_us": round(float(np.median(x)), 2), "p95_us": round(float(np.percentile(x, 95)), 2),
            "max_us": round(float(x.max()), 2)}


out["cta_total"] = stats(tend - t0)
out["stage_map_and_barrier_init"] = stats(tmap - t0)
out["role_setup_and_link_load"] = stats(tloop0 - tmap)
out["eigenvector_loop"] = stats(tloop1 - tloop0)
out["epilogue"] = stats(tend - tloop1)
ck = t[:, 8:14].astype(np.float64)
out["cycles"] = {"stage_map": round(float((ck[:, 2] - ck[:, 1]).mean())), "role_setup": round(float((ck[:, 3] - ck[:, 2]).mean())),
                 "loop": round(float((ck[:, 4] - ck[:, 3]).mean())), "epilogue": round(float((ck[:, 5] - ck[:, 4]).mean())),
                 "loop_per_eigenvector": round(float((ck[:, 4] - ck[:, 3]).mean()) / nev, 1)}
out["per_eigenvector_us"] = round(float((tloop1 - tloop0).mean()) / 1e3 / nev, 3)
gaps, busy, last_end = [], [], []
for s in np.unique(sm):
    idx = np.where(sm == s)[0]
    o = idx[np.argsort(t0[idx])]
    if len(o) > 1:
        gaps.extend((t0[o][1:] - tend[o][:-1]).tolist())
    busy.append(float((tend[o] - t0[o]).sum()))
    last_end.append(float(tend[o].max()))
out["gap_between_ctas_on_an_sm"] = stats(gaps) if gaps else None
span = float(tend.max() - start)
out["sm_busy_fraction"] = round(float(np.mean(busy)) / span, 4)
out["loop_fraction_of_span"] = round(float((tloop1 - tloop0).sum()) / len(np.unique(sm)) / span, 4)
out["tail_idle_fraction"] = round(float(np.mean(tend.max() - np.array(last_end))) / span, 4)
out["sms_used"] = int(len(np.unique(sm)))
print(json.dumps(out))
