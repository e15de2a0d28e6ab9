"""Per-stage kernel rooflines (SURVEY.md §8a rows a3, a6, a8, a10 as separate C-ABI calls): each reference-shaped entry
point of include/mugiq_b200.h timed alone with CUDA events (torch's current stream = the launching stream), L2 flushed
between timed launches unless the operands are larger than L2 anyway, reported as algorithmic GB/s against the
measured HBM peak (MEASURED_PEAKS.json) and, for the projections, FP64 TFLOP/s against the measured DMMA peak.

    python tools/stage_bench.py [--out gpurun_out/stage_bench.json] [--big]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mugiq_b200 import ops, synth  # noqa: E402
from mugiq_b200.params import momenta_up_to  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
S, U, A = 192, 144, 256


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6551.4


_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def timeit(fn, reps=10, warm=3, flush=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush:
            flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def row(name, L, ms, best, nbytes, flops=0.0, note=""):
    hbm = peaks()
    r = {"kernel": name, "L": list(L), "ms_median": round(ms, 5), "ms_best": round(best, 5),
         "alg_bytes": nbytes, "GBps": round(nbytes / ms / 1e6, 1), "frac_hbm_peak": round(nbytes / ms / 1e6 / hbm, 4)}
    if flops:
        r["TFLOPs"] = round(flops / ms / 1e9, 2)
        r["frac_dmma_peak"] = round(flops / ms / 1e9 / 37.1, 4)
    if note:
        r["note"] = note
    print(json.dumps(r), flush=True)
    return r


def randc(*shape):
    return torch.randn(*shape, dtype=torch.complex128, device="cuda")


def bench_lattice(L, nvec, out):
    V4 = L[0] * L[1] * L[2] * L[3]
    ev = randc(nvec, V4, 12)
    ev2 = randc(nvec, V4, 12)
    gauge = randc(4, V4, 3, 3)
    loop = torch.zeros(16, V4, dtype=torch.complex128, device="cuda")
    sig = [0.01 + 0.001 * i for i in range(nvec)]
    vl = [ev[i] for i in range(nvec)]
    vr = [ev2[i] for i in range(nvec)]
    big = nvec * V4 * S > (256 << 20)

    # a3 contraction
    ms, b = timeit(lambda: ops.contract_batch(loop, vl, None, sig, L, accumulate=False), flush=not big)
    out.append(row(f"contract_batch ultra-local x{nvec}", L, ms, b, V4 * (nvec * S + A)))
    ms, b = timeit(lambda: ops.contract_batch(loop, vl, vr, sig, L, accumulate=False), flush=not big)
    out.append(row(f"contract_batch vL!=vR x{nvec}", L, ms, b, V4 * (nvec * 2 * S + A)))
    ms, b = timeit(lambda: ops.contract(loop, ev[0], ev2[0], 0.5, L))
    out.append(row("contract single pair (performLoopContraction)", L, ms, b, V4 * (2 * S + 2 * A)))

    # a6 displacement
    for d, s in ((0, 1), (0, 0), (1, 1), (2, 0), (3, 1), (3, 0)):
        ms, b = timeit(lambda: ops.displace(ev2[0], ev[0], gauge, d, s, L))
        out.append(row(f"displace single dir={d} sign={s} (performCovariantDisplacementVector)", L, ms, b, V4 * (2 * S + U)))
    nb = min(nvec, 32)
    for d, s in ((0, 1), (1, 0), (3, 1)):
        ms, b = timeit(lambda: ops.displace_batch(vr[:nb], vl[:nb], gauge, d, s, L), flush=not big)
        out.append(row(f"displace_batch x{nb} dir={d} sign={s}", L, ms, b, V4 * (nb * 2 * S + U)))
    # the same two stages on fields in QUDA FLOAT2 order (thread = site, coalesced without staging)
    ms, b = timeit(lambda: ops.contract_native(loop, vl, vr, sig, 2, L, accumulate=False), flush=not big)
    out.append(row(f"contract_native FLOAT2 vL!=vR x{nvec}", L, ms, b, V4 * (nvec * 2 * S + A)))
    ms, b = timeit(lambda: ops.contract_native(loop, vl[:1], vr[:1], sig[:1], 2, L, accumulate=True))
    out.append(row("contract_native FLOAT2 single pair", L, ms, b, V4 * (2 * S + 2 * A)))
    nbn = min(nvec, 16)
    for d, s in ((0, 1), (3, 0)):
        ms, b = timeit(lambda: ops.displace_native(vr[:nbn], vl[:nbn], gauge, d, s, 2, L), flush=not big)
        out.append(row(f"displace_native FLOAT2 x{nbn} dir={d} sign={s}", L, ms, b, V4 * (nbn * 2 * S + U)))
        ms, b = timeit(lambda: ops.displace_native(vr[:1], vl[:1], gauge, d, s, 2, L))
        out.append(row(f"displace_native FLOAT2 single dir={d} sign={s}", L, ms, b, V4 * (2 * S + U)))
    del ev2, vr

    # a8 reorder
    nLoop = 9
    pos = randc(nLoop, 16, V4)
    mp = torch.empty(V4 // L[3] * 16 * nLoop * L[3], dtype=torch.complex128, device="cuda")
    ms, b = timeit(lambda: ops.reorder_mapgamma(mp, pos, 16 * nLoop, nLoop, L), flush=False)
    out.append(row(f"reorder_mapgamma nLoop={nLoop}", L, ms, b, 2 * 16 * 16 * V4 * nLoop))
    del pos, mp


def bench_projection(L, nLoop, pmax2, out):
    V4 = L[0] * L[1] * L[2] * L[3]
    V3 = V4 // L[3]
    mom = momenta_up_to(pmax2)
    N = len(mom)
    pos = randc(nLoop, 16, V4)
    ph = ops.phase_matrix_eo(mom, -1, L)
    M = L[3] * 16 * nLoop
    flops = 8.0 * M * N * V3
    nbytes = 16.0 * (M * V3 + N * V3 + M * N)
    ws = torch.empty(max(ops.momproj_pos_workspace_bytes(L, 8, nLoop, N), 16), dtype=torch.uint8, device="cuda")
    ms, b = timeit(lambda: ops.momproj_pos(pos, ph, nLoop, L, workspace=ws), flush=False)
    out.append(row(f"momproj_pos (stages 3+4) nLoop={nLoop} Nmom={N}", L, ms, b, nbytes, flops))
    # two-call form: reorder + GEMM
    mp = torch.empty(M * V3, dtype=torch.complex128, device="cuda")
    ops.reorder_mapgamma(mp, pos, 16 * nLoop, nLoop, L)
    del pos
    ph2 = ops.phase_matrix(mom, -1, L)
    ms, b = timeit(lambda: ops.momproj(mp, ph2, M, N, V3), flush=False)
    out.append(row(f"momproj GEMM (stage 4) M={M} N={N} K={V3}", L, ms, b, nbytes, flops))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/stage_bench.json")
    ap.add_argument("--big", action="store_true", help="also 32^3x64 and config 3's projection")
    a = ap.parse_args()
    res = []
    bench_lattice((16, 16, 16, 32), 64, res)
    bench_projection((16, 16, 16, 32), 9, 1, res)
    if a.big:
        bench_lattice((32, 32, 32, 64), 16, res)
        bench_lattice((24, 24, 24, 48), 16, res)
        bench_projection((24, 24, 24, 48), 33, 4, res)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump({"hbm_peak_GBps": peaks(), "dmma_peak_TFLOPs": 37.1, "rows": res}, open(a.out, "w"), indent=1)
