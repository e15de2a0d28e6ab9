#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the disconnected-loop hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

One "step" = one pass of the whole hot path (stage 1 contraction over the 16 gammas + stage 2 covariant
displacements + stage 3 gamma-map/time-slice reorder + stage 4 momentum projection) over the full eigenvector
set of the workload.  Metric (BASELINE.json): eigvec·site contractions/s (16 gamma, all displacements)
= nEv * V4 * nLoop / step time.  At N > 1 every rank holds its own shard of `nev` eigenvectors (weak scaling,
gauge field replicated) and the step ends with the NCCL all-reduce of the loop buffer.

Prints ONE JSON line on rank 0 (see the task contract): value (inputs resident in HBM), e2e (host buffers through
the public Loop_Mugiq API, H2D/D2H inside the timed region), roofline of the dominant kernel measured live with
CUDA events on the launching stream, cpu_baseline (the oracle port on the host cores, bounded sample), clocks.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "eigvec_site_contractions_per_s"
UNIT = "eigvec*site*loop contractions/s (16 gamma each)"

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on; fits one GPU (5.0 GB of eigenvectors)
    "16x16x16x32_nev200_ulocal+1hop8": dict(L=(16, 16, 16, 32), nev=200, entries="+x:1;-x:1;+y:1;-y:1;+z:1;-z:1;+t:1;-t:1",
                                           p2max=1),
    # BASELINE.json configs[2]
    "24x24x24x48_nev500_disp1to4_p2le4": dict(L=(24, 24, 24, 48), nev=500,
                                              entries="+x:1,4;-x:1,4;+y:1,4;-y:1,4;+z:1,4;-z:1,4;+t:1,4;-t:1,4", p2max=4),
    # BASELINE.json configs[3], per-GPU share (1000 eigenvectors over 8 GPUs), ultra-local only
    "32x32x32x64_nev125_ulocal": dict(L=(32, 32, 32, 64), nev=125, entries="", p2max=0),
    # BASELINE.json configs[4], per-GPU time slab of the 48^3x96 lattice on 8 GPUs (use with --tsplit --gpus 8): 300 of the
    # 2000 eigenvectors are resident (102 GB of extended slabs per GPU; the full set has to stream from the host)
    "48x48x48x12_nev300_ulocal+1hop8": dict(L=(48, 48, 48, 12), nev=300, entries="+x:1;-x:1;+y:1;-y:1;+z:1;-z:1;+t:1;-t:1",
                                           p2max=1),
    # small case for quick checks
    "8x8x8x16_nev16_ulocal+1hop8": dict(L=(8, 8, 8, 16), nev=16, entries="+x:1;-x:1;+y:1;-y:1;+z:1;-z:1;+t:1;-t:1", p2max=1),
}
DEFAULT_WORKLOAD = "16x16x16x32_nev200_ulocal+1hop8"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # NVML missing: report nulls
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.004)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_device_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except (ValueError, IndexError):
            return local_rank
    return local_rank


def loop_count(entries):
    return 1 + sum(b - a + 1 for (_, _, a, b) in entries)


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port (oracle/mugiq_oracle.cpp) on the host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_loop_rate(wl, nev_sample, reps=1):
    """Times the oracle's Loop_Mugiq::computeCoarseLoop restatement on `nev_sample` eigenvectors of the workload
    (same lattice, same displacement entries), all host threads.  Returns (contractions/s, threads, seconds)."""
    from oracle import oracle as orc
    from mugiq_b200 import synth
    from mugiq_b200.params import parse_disp_entries, which_displace
    L = wl["L"]
    entries = []
    if wl["entries"]:
        _, ds, a, b = parse_disp_entries(wl["entries"])
        entries = [which_displace(s) + (x, y) for s, x, y in zip(ds, a, b)]
    ev = synth.random_evecs_np(L, nev_sample, seed=5)
    U = synth.random_gauge(L, seed=5) if entries else None
    sig = synth.sigmas(nev_sample)
    orc.compute_loop(ev[:1], sig[:1], U, entries[:1], L)  # warm-up (page faults, thread pool)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        orc.compute_loop(ev, sig, U, entries, L)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    V4 = int(np.prod(L))
    return nev_sample * V4 * loop_count(entries) / best, orc.num_threads(), best


def cpu_sample_size(wl, target_s):
    """Eigenvectors per CPU step so that one step costs about `target_s` seconds: calibrated with a 2-eigenvector
    run on this host (the GPU box's core count differs from the build container's)."""
    _, _, dt2 = cpu_loop_rate(wl, 2)
    return int(max(2, min(wl["nev"], 2 * target_s / dt2)))


def reference_gpu_rate(wl, nev_sample):
    """The reference's OWN CUDA kernels and wrappers (oracle/_ref/libmugiq_ref.so: lib/contract_wrappers.cu +
    lib/mugiq_*_kernels.cu compiled unmodified against oracle/quda_shim) in the reference's loop order
    (lib/loop_mugiq.cpp:455-509 restated in oracle/ref_driver.cu) on this GPU, `nev_sample` eigenvectors of the workload.
    A reported baseline (what the reference's GPU path does on a B200), not the product path.  None if the library was
    not built (it needs /root/reference at build time) or no GPU is visible."""
    try:
        import torch
        from oracle import ref_kernels as ref
        if not (ref.available() and torch.cuda.is_available()):
            return None
        from mugiq_b200.params import parse_disp_entries, which_displace
        L = wl["L"]
        entries = []
        if wl["entries"]:
            _, ds, a, b = parse_disp_entries(wl["entries"])
            entries = [which_displace(s) + (x, y) for s, x, y in zip(ds, a, b)]
        V4 = int(np.prod(L))
        evq = [torch.randn(V4, 12, dtype=torch.complex128, device="cuda") for _ in range(nev_sample)]  # QUDA FLOAT2 order
        gauge = torch.randn(4, V4, 3, 3, dtype=torch.complex128, device="cuda")
        sig = [0.01 + 0.001 * i for i in range(nev_sample)]
        ref.compute_loop(evq[:1], sig[:1], gauge, entries, L)  # warm-up: module load, constant tables
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ref.compute_loop(evq, sig, gauge, entries, L)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return {"value": nev_sample * V4 * loop_count(entries) / dt, "unit": UNIT, "kind": "reference kernels on this GPU",
                "seconds": dt,
                "sample": f"{nev_sample} of {wl['nev']} eigenvectors, full lattice, all {loop_count(entries)} loops; "
                          "lib/contract_wrappers.cu + lib/mugiq_*_kernels.cu unmodified (oracle/_ref), loop nest of "
                          "lib/loop_mugiq.cpp:455-509 restated around them, device printf of the contraction kernel sent to /dev/null"}
    except Exception as exc:  # a baseline must never take the bench down
        return {"unavailable": repr(exc)[:200]}


def run_reference(args, wl, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as orc
    orc.set_num_threads(len(os.sched_getaffinity(0)))  # torchrun exports OMP_NUM_THREADS=1: use every host core anyway
    nev_s = cpu_sample_size(wl, min(15.0, 150.0 / (args.steps + 1)))
    times = []
    rate = threads = None
    for i in range(args.warmup + args.steps):
        if i < args.warmup and i > 0:
            continue  # one warm-up pass is enough for a CPU loop
        r, threads, dt = cpu_loop_rate(wl, nev_s)
        if i >= args.warmup:
            times.append(dt)
            rate = r if rate is None else max(rate, r)
    V4 = int(np.prod(wl["L"]))
    from mugiq_b200.params import parse_disp_entries
    nloop = 1 + (sum(y - x + 1 for x, y in zip(*parse_disp_entries(wl["entries"])[2:])) if wl["entries"] else 0)
    mean_dt = float(np.mean(times))
    value = nev_s * V4 * nloop / mean_dt
    sample = f"{nev_s} of {wl['nev']} eigenvectors per step, full lattice, all loops (oracle port, OpenMP)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": mean_dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "L": list(wl["L"]), "nev": wl["nev"], "entries": wl["entries"], "nLoop": nloop},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "reference_gpu": reference_gpu_rate(wl, min(wl["nev"], 20))}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args, wl, name):
    import torch
    import torch.distributed as dist
    from mugiq_b200 import ops, synth
    from mugiq_b200.loop import Loop_Mugiq, Eigsolve
    from mugiq_b200.params import MugiqLoopParam, momenta_up_to

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        # NCCL_DEBUG=VERSION makes NCCL print its version on stdout, which must carry exactly one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    L, nev = wl["L"], args.nev or wl["nev"]
    V4 = int(np.prod(L))
    V3 = V4 // L[3]
    ts = None
    if args.tsplit:
        # lattice-T split: every rank owns a slab of wl["L"] of a global lattice with T = L[3] * world; the eigenvectors
        # span all ranks (same sigma everywhere), the halos of +-t displacements travel over NVLink (NCCL P2P)
        from mugiq_b200.tsplit import TSplit
        from mugiq_b200.params import parse_disp_entries
        tmax = 0
        if wl["entries"]:
            _, ds, _, stop = parse_disp_entries(wl["entries"])
            tmax = max([b for s, b in zip(ds, stop) if s[1] == "t"] + [0])
        Lg = (L[0], L[1], L[2], L[3] * world)
        ts = TSplit(Lg, rank, world, tmax)
        # this rank's extended slab of one global field, generated on the device slice by slice (slice-keyed seeds)
        U = synth.random_gauge_slab_torch(Lg, [(ts.t0 - ts.H + i) % Lg[3] for i in range(ts.Tl + 2 * ts.H)], seed=11,
                                          device=torch.device("cuda", local_rank))
    else:
        U = synth.random_gauge(L, seed=11)
    prm = MugiqLoopParam(gauge=[U[mu] for mu in range(4)])
    if wl["entries"]:
        prm.set_displacements(wl["entries"])
    mom = momenta_up_to(wl["p2max"])
    prm.set_momenta(mom)
    sig = synth.sigmas(nev) + (0.0 if ts is not None else 0.2 * rank)
    if ts is not None:
        # the slab is stored in the extended layout (halo slices allocated, filled by the exchange every step)
        if args.halo == "nccl":
            ev_d = synth.random_evecs_torch(ts.L_ext, nev, seed=100 + rank, device=dev)
        else:
            # NVLink peer mode: the slabs live in one IPC-shareable allocation the two time neighbours map; halos are
            # written straight into the neighbours' slabs (copy engines or SM kernel)
            V4e = int(np.prod(ts.L_ext))
            peer_buf = ops.PeerBuffer(nev * V4e * 12 * 16, device=dev)
            ev_d = synth.random_evecs_torch(ts.L_ext, nev, seed=100 + rank, device=dev,
                                            out=peer_buf.tensor((nev, V4e, 12), torch.complex128))
            if world > 1:
                handles = [None] * world
                dist.all_gather_object(handles, peer_buf.handle)
                up_ptr = ops.peer_open(handles[(rank + 1) % world], dev)
                dn_ptr = up_ptr if world == 2 else ops.peer_open(handles[(rank - 1) % world], dev)
            else:
                up_ptr = dn_ptr = peer_buf.ptr
            ts.attach_peers(ev_d, up_ptr, dn_ptr, mode=0 if args.halo == "dma" else 1, group=group)
        es = Eigsolve(list(ev_d), sig, L, ext_volume=int(np.prod(ts.L_ext)))
    else:
        ev_d = synth.random_evecs_torch(L, nev, seed=100 + rank, device=dev)      # [nev, V4, 12] resident in HBM
        es = Eigsolve(list(ev_d), sig, L)
    loop = Loop_Mugiq(prm, es, device=dev, group=group, evec_batch=args.evec_batch,
                      copy_pos_to_host=False, tsplit=ts, stream_batch=args.tsplit_batch if ts is not None else 16)
    nLoop = loop.cPrm.nLoop
    units_per_rank = nev * V4 * nLoop

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        barrier()
        ops.prof_reset()
        ops.prof_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler is not None:
            sampler.__enter__()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        if sampler is not None:
            sampler.__exit__()
        ops.prof_enable(False)
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    # ---- leg 1: inputs resident in HBM --------------------------------------------------------------------
    def step_resident():
        loop.MomProjDone = False
        loop.computeCoarseLoop()

    sampler = ClockSampler(physical_device_index(local_rank))
    ms_step = timed(step_resident, args.steps, args.warmup, sampler)
    halo_sides = getattr(loop, "tsplit_halo_sides", 2)
    halo_slices = getattr(loop, "tsplit_halo_slices", 2 * ts.H if ts is not None else 0)
    if ts is not None and getattr(loop, "_trace_on", False):
        torch.cuda.synchronize()
        print(f"[rank {rank}] T-split phases of the last step (device ms, host ms): " +
              "; ".join(f"{n} {g:.3f}/{h:.3f}" for n, g, h in loop.trace_report()), file=sys.stderr, flush=True)
    report = ops.prof_report()
    launches = sum(v["launches"] for v in report.values())
    value = world * units_per_rank / (ms_step * 1e-3)

    # dominant kernel and its roofline
    peaks, peak_src = measured_peaks()
    dom = max(report, key=lambda k: report[k]["ms"]) if report else None
    roofline = None
    if dom is not None and report[dom]["timed"] > 0:
        d = report[dom]
        per_launch_ms = d["ms"] / d["timed"]
        per_launch_bytes = d["alg_bytes"] / d["launches"]
        achieved = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get(dom, {}).get("dram_bytes_per_launch")
        fp64 = None
        if d.get("alg_flops", 0) > 0:
            mb = os.path.join(ROOT, "profiles", "r1_microbench_b200.json")
            fp64_peak = json.load(open(mb))["dfma_tflops"] if os.path.exists(mb) else 34.15
            tf = d["alg_flops"] / d["launches"] / (per_launch_ms * 1e-3) / 1e12
            fp64 = {"achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
                    "peak_source": "measured DFMA stream, tools/microbench.cu (profiles/r1_microbench_b200.json)",
                    "note": "the fused kernel is FP64-pipe bound (15 flop per compulsory byte, ridge 5.2): this fraction, "
                            "not the HBM one, measures its distance from the speed of light (DESIGN.md 4.1)"}
        roofline = {"bound": "hbm", "kernel": dom, "fp64": fp64, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_src,
                    "alg_bytes_per_launch": per_launch_bytes, "ms_per_launch": per_launch_ms,
                    "share_of_step": d["ms"] / (ms_step * args.steps),
                    "kernels": {k: {"launches": v["launches"], "ms": round(v["ms"], 4),
                                    "GBps": (v["alg_bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 else None}
                                for k, v in report.items()}}
        # stage 4 (momentum projection) separately, as SURVEY 8d asks: useful FP64 TFLOP/s (8*M*N*K) against the measured
        # DMMA peak; with few momenta (N <~ 12) the projection is HBM bound and the GB/s figure above is the relevant one
        mp = report.get("momproj")
        if mp and mp["ms"] > 0 and mp.get("alg_flops", 0) > 0:
            tf = mp["alg_flops"] / (mp["ms"] * 1e-3) / 1e12
            roofline["projection"] = {"TFLOPs": tf, "dmma_peak": 37.1, "frac": tf / 37.1, "Nmom": len(mom),
                                      "bound": "tensor" if len(mom) >= 12 else "hbm",
                                      "peak_source": "measured DMMA stream, tools/microbench.cu"}

    # ---- leg 2: end to end through the public API with HOST buffers -----------------------------------------
    e2e = None
    if not args.no_e2e and ts is None:
        ev_h = torch.empty((nev, V4, 12), dtype=torch.complex128, pin_memory=True)
        ev_h.copy_(ev_d)
        del loop
        loop_h = Loop_Mugiq(prm, Eigsolve(list(ev_h), sig, L), device=dev, group=group, evec_batch=args.evec_batch,
                            copy_pos_to_host=True)

        def step_host():
            loop_h.MomProjDone = False
            loop_h.displace.upload_gauge(prm) if loop_h.displace is not None else None   # H2D of the gauge field
            loop_h.computeCoarseLoop()                                                    # H2D evecs, D2H dataPos + dataMom

        e2e_steps = max(1, min(args.steps, 5))
        ms_e2e = timed(step_host, e2e_steps, 1)
        h2d = ev_h.numel() * 16 + (U.nbytes if loop_h.displace is not None else 0)
        d2h = loop_h.dataPos.numel() * 16 + loop_h.dataMom.numel() * 16
        e2e = {"value": world * units_per_rank / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e, "steps": e2e_steps}

    # ---- leg 3: eigenvectors resident in QUDA's native FLOAT2 order (what a QUDA-backed caller holds) ----------------
    # the layout conversion into the canonical site-major order (one batched launch) is inside the timed region
    quda = None
    if not args.no_e2e and ts is None and world == 1:
        if e2e is not None:
            del loop_h
        ev_q = [ops.export_spinor(ev_d[i], 2, L) for i in range(nev)]
        stage = torch.empty_like(ev_d)
        loop_q = Loop_Mugiq(prm, Eigsolve(list(stage), sig, L), device=dev, group=group, evec_batch=args.evec_batch,
                            copy_pos_to_host=False)

        def step_quda():
            ops.ingest_spinor_batch(ev_q, 2, L, out=stage)
            loop_q.MomProjDone = False
            loop_q.computeCoarseLoop()

        q_steps = max(1, min(args.steps, 5))
        ms_q = timed(step_quda, q_steps, 2)
        quda = {"value": world * units_per_rank / (ms_q * 1e-3), "unit": UNIT, "ms_per_step": ms_q, "steps": q_steps,
                "note": "eigenvectors resident in HBM in QUDA FLOAT2 order; mugiq_b200_ingest_spinor_batch (FLOAT2 -> site-major, "
                        "2 x 192 B per eigvec*site) runs inside every step"}
        del loop_q, stage, ev_q

    # ---- cpu baseline (rank 0, N == 1 only) ---------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as orc
        orc.set_num_threads(len(os.sched_getaffinity(0)))
        nev_s = cpu_sample_size(wl, 12.0)
        rate, threads, dt = cpu_loop_rate(wl, nev_s)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{nev_s} of {nev} eigenvectors, full lattice, all {nLoop} loops, {dt:.1f} s (oracle port, OpenMP)"}

    ref_gpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        if e2e is not None:
            del ev_h
            if quda is None:
                del loop_h
        del ev_d, es
        torch.cuda.empty_cache()
        ref_gpu = reference_gpu_rate(wl, min(nev, 20))

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": name, "L": list(L), "nev_per_gpu": nev, "entries": wl["entries"], "nLoop": nLoop,
                           "Nmom": len(mom), "stages": "contract+displace+reorder+momproj" + (("+halo exchange+allgather" if ts is not None else "+allreduce") if world > 1 else ""),
                           "l2": f"inputs larger than L2 ({nev * V4 * 192 / 1e9:.2f} GB of eigenvectors read per step)",
                           "evec_batch": args.evec_batch,
                           "partition": ("lattice-T split, global T = %d, halo %d slices, %.1f MB of halo per rank and step over "
                                         "NVLink (%d-sided eigenvector halo of %d slice(s), transport %s; interior-only compute)"
                                         % (L[3] * world, ts.H, nev * ts.halo_bytes_per_vector(slices=halo_slices) / 1e6, halo_sides,
                                            halo_slices, args.halo))
                           if ts is not None else ("eigenvector shards" if world > 1 else "single GPU")},
                "roofline": roofline, "cpu_baseline": cpu, "reference_gpu": ref_gpu, "e2e": e2e, "quda_order": quda, "gpu_launches": launches,
                "clocks": sampler.summary()}
        print(json.dumps(line), flush=True)
    if ts is not None and ts.peer is not None:
        # unmap the neighbours' slabs before anybody frees its own
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            for ptr in {up_ptr, dn_ptr}:
                ops.peer_close(ptr, dev)
            dist.barrier()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS))
    ap.add_argument("--nev", type=int, default=0, help="override the eigenvector count per GPU (debugging)")
    ap.add_argument("--evec-batch", type=int, default=200)
    ap.add_argument("--tsplit", action="store_true", help="partition the lattice in T over the GPUs (halo exchange) instead of "
                                                          "sharding eigenvectors")
    ap.add_argument("--halo", default="dma", choices=["dma", "kernel", "nccl"],
                    help="--tsplit halo transport: direct NVLink writes into the neighbours' slabs by the copy engines (dma) or "
                         "an SM push kernel (kernel), or NCCL send/recv through staging buffers (nccl)")
    ap.add_argument("--tsplit-batch", type=int, default=100, help="--tsplit: eigenvectors per halo push / kernel launch (the push "
                                                                  "of batch i+1 overlaps the kernels of batch i)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl, args.workload)
    return run_ours(args, wl, args.workload)


if __name__ == "__main__":
    sys.exit(main())
