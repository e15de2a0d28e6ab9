"""Bandwidth of the QUDA-native <-> site-major layout conversion (one field per launch and 32 per launch)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mugiq_b200 import ops, synth
L = (16, 16, 16, 32)
v = synth.random_evecs_torch(L, 32, seed=1)
for order in (2, 4):
    q = [ops.export_spinor(v[k], order, L) for k in range(32)]
    ops.prof_reset(); ops.prof_enable(True)
    for _ in range(5):
        for k in range(32):
            s = ops.ingest_spinor(q[k], order, L)
    r = ops.prof_report()["convert_spinor"]
    one = r["alg_bytes"] / r["ms"] / 1e6
    ops.prof_reset()
    for _ in range(5):
        b = ops.ingest_spinor_batch(q, order, L)
    r = ops.prof_report()["convert_spinor"]
    assert torch.equal(b, v) and torch.equal(s, v[31])
    print(f"order FLOAT{order}: single {one:.0f} GB/s, batch of 32 {r['alg_bytes'] / r['ms'] / 1e6:.0f} GB/s")
