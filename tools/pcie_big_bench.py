"""H2D rate of ONE pinned buffer as large as a step's eigenvector set (5 GB) against the 1 GiB figure, whole and in 400 MB
pieces; D2H of 302 MB (dataPos).  One JSON object."""
import json

import torch

out = {}
for gib in (1, 5):
    n = gib << 30
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for pieces in (1, 13):
        best = 1e30
        step = n // pieces
        for _ in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(pieces):
                d[k * step:(k + 1) * step].copy_(h[k * step:(k + 1) * step], non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[f"h2d_{gib}GiB_{pieces}pieces_GBps"] = step * pieces / best / 1e6
    del h, d
m = 302505984
h = torch.empty(m, dtype=torch.uint8, pin_memory=True)
d = torch.empty(m, dtype=torch.uint8, device="cuda")
best = 1e30
for _ in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); h.copy_(d, non_blocking=True); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
out["d2h_302MB_GBps"] = m / best / 1e6
out["d2h_302MB_ms"] = best
print(json.dumps(out))
