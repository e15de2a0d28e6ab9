"""Host<->device copy bandwidth of the box (pinned memory), the ceiling of bench.py's e2e leg."""
import torch, json
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
out = {}
for name, (dst, src) in {"h2d": (d, h), "d2h": (h, d)}.items():
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dst.copy_(src, non_blocking=True); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[name + "_GBps_1GiB"] = n / best / 1e6
# 25 MB pieces (one eigenvector of 16^3x32), back to back
p = 25165824
best = 1e30
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(40):
        d[k * p:(k + 1) * p].copy_(h[k * p:(k + 1) * p], non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
out["h2d_GBps_40x25MB"] = 40 * p / best / 1e6
print(json.dumps(out))
