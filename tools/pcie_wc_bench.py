"""Does write-combined pinned host memory (cudaHostAllocWriteCombined) or splitting a copy over two streams move the H2D
ceiling of bench.py's e2e leg?  Prints one JSON object (GB/s, 1 GiB copies, best of 5)."""
import ctypes as C
import json

import torch

rt = C.CDLL("libcudart.so.12")
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
out = {}


def time_copy(hptr, nstreams=1):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    best = 1e30
    part = n // nstreams
    for _ in range(5):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        ends = []
        for k, s in enumerate(streams):
            s.wait_event(e0)
            rt.cudaMemcpyAsync(C.c_void_p(d.data_ptr() + k * part), C.c_void_p(hptr + k * part), C.c_size_t(part), 1, C.c_void_p(s.cuda_stream))
            e = torch.cuda.Event(enable_timing=True)
            e.record(s)
            ends.append(e)
        torch.cuda.synchronize()
        best = min(best, max(e0.elapsed_time(e) for e in ends))
    return n / best / 1e6


for name, flags in (("pinned_default", 0), ("pinned_write_combined", 4)):
    p = C.c_void_p()
    rc = rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags))
    if rc:
        out[name] = f"cudaHostAlloc failed: {rc}"
        continue
    C.memset(p, 1, n)
    out[name + "_h2d_GBps"] = time_copy(p.value)
    out[name + "_h2d_GBps_2streams"] = time_copy(p.value, 2)
    rt.cudaFreeHost(p)
print(json.dumps(out))
