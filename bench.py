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
CUDA events on the launching stream, cpu_baseline (the oracle port on the host cores, bounded sample), clocks,
`verified` (a 4-eigenvector subset through the same path against the oracle + a checksum of the timed buffer).
Extra objects: `config4` (BASELINE configs[3]: 32^3x64, 125 eigenvectors per GPU, ultra-local + 8 one-hop loops, the
position-space loop buffer summed over the GPUs with the chunked all-reduce that overlaps the kernels), and at N > 1
`tsplit` (BASELINE configs[4]: 48^3 time slabs, halos over NVLink peer memory).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "eigvec_site_contractions_per_s"
UNIT = "eigvec*site*loop contractions/s (16 gamma each)"
ONE_HOP = "+x:1;-x:1;+y:1;-y:1;+z:1;-z:1;+t:1;-t:1"

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on; fits one GPU (5.0 GB of eigenvectors)
    "16x16x16x32_nev200_ulocal+1hop8": dict(L=(16, 16, 16, 32), nev=200, entries=ONE_HOP, p2max=1),
    # BASELINE.json configs[2]
    "24x24x24x48_nev500_disp1to4_p2le4": dict(L=(24, 24, 24, 48), nev=500,
                                              entries="+x:1,4;-x:1,4;+y:1,4;-y:1,4;+z:1,4;-z:1,4;+t:1,4;-t:1,4", p2max=4),
    # BASELINE.json configs[3], per-GPU share (1000 eigenvectors over 8 GPUs): ultra-local only / with the one-hop loops
    "32x32x32x64_nev125_ulocal": dict(L=(32, 32, 32, 64), nev=125, entries="", p2max=0),
    "32x32x32x64_nev125_ulocal+1hop8": dict(L=(32, 32, 32, 64), nev=125, entries=ONE_HOP, p2max=1),
    # BASELINE.json configs[4], per-GPU time slab of the 48^3x96 lattice on 8 GPUs (use with --tsplit --gpus 8): 300 of the
    # 2000 eigenvectors are resident (102 GB of extended slabs per GPU; the full set has to stream from the host)
    "48x48x48x12_nev300_ulocal+1hop8": dict(L=(48, 48, 48, 12), nev=300, entries=ONE_HOP, p2max=1),
    # small case for quick checks
    "8x8x8x16_nev16_ulocal+1hop8": dict(L=(8, 8, 8, 16), nev=16, entries=ONE_HOP, p2max=1),
}
DEFAULT_WORKLOAD = "16x16x16x32_nev200_ulocal+1hop8"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def fp64_peaks():
    """FP64 denominators measured on this pool's B200s by tools/microbench.cu (DFMA stream, DMMA stream) and
    tools/fp64_gemm_peak.py (cuBLAS ZGEMM 4096^3): the driver's MEASURED_PEAKS.json has no FP64 figure."""
    out = {"dfma_tflops": 34.03, "dmma_tflops": 37.06, "zgemm_4096_tflops": 36.8, "source": "built-in copy of profiles/r1_*"}
    try:
        with open(os.path.join(ROOT, "profiles", "r1_microbench_b200.json")) as fh:
            mb = json.load(fh)
        with open(os.path.join(ROOT, "profiles", "r1_fp64_gemm_peak.json")) as fh:
            gp = json.load(fh)
        out = {"dfma_tflops": mb["dfma_tflops"], "dmma_tflops": mb["dmma_tflops"], "zgemm_4096_tflops": gp["zgemm_4096_tflops"],
               "source": "measured: tools/microbench.cu (profiles/r1_microbench_b200.json), cuBLAS ZGEMM 4096^3 "
                         "(profiles/r1_fp64_gemm_peak.json)"}
    except Exception:
        pass
    # round 3: cycles per DFMA of the operand patterns (tools/dfma_bench.cu).  The fused kernel's FMAs are complex 4x4
    # outer products (two fresh 64-bit register operands per instruction): that stream tops out below the chained one
    try:
        with open(os.path.join(ROOT, "profiles", "r3_dfma_bench.json")) as fh:
            db = json.load(fh)
        out["dfma_cplx_outer_tflops"] = db["cplx_outer4x4_8warps"]["tflops"]
        out["dfma_cplx_outer_sustained_tflops"] = db.get("cplx_outer4x4_sustained_6s", {}).get("tflops")
        out["dfma_cplx_outer_cycles"] = db["cplx_outer4x4_8warps"]["cycles_per_dfma_per_smsp"]
        out["dfma_chain_cycles"] = db["chain_8warps"]["cycles_per_dfma_per_smsp"]
    except Exception:
        pass
    return out


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


def bind_near_gpu(index):
    """Pins this process to the CPU cores NVML reports as local to its GPU, so that the pinned host buffers of the
    end-to-end leg are first-touched on the GPU's own NUMA node (8 ranks share the host's memory fabric).  Returns the
    number of cores bound to, or None when NVML / the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def loop_count(entries):
    return 1 + sum(b - a + 1 for (_, _, a, b) in entries)


def entries_of(wl):
    from mugiq_b200.params import parse_disp_entries, which_displace
    if not wl["entries"]:
        return []
    _, ds, a, b = parse_disp_entries(wl["entries"])
    return [which_displace(s) + (x, y) for s, x, y in zip(ds, a, b)]


def make_config(name, wl, nev, world, evec_batch, partition=None):
    """The `config` object of the JSON line - built by ONE function for both arms, so that the driver sees the same
    workload description from this repo's arm and from the reference arm."""
    from mugiq_b200.params import momenta_up_to
    L = wl["L"]
    V4 = int(np.prod(L))
    return {"workload": name, "L": list(L), "nev_per_gpu": nev, "entries": wl["entries"], "nLoop": loop_count(entries_of(wl)),
            "Nmom": len(momenta_up_to(wl["p2max"])),
            "stages": "contract+displace+reorder+momproj" + ("" if world == 1 else "+cross-GPU sum"),
            "l2": f"inputs larger than L2 ({nev * V4 * 192 / 1e9:.2f} GB of eigenvectors read per step)",
            "evec_batch": evec_batch,
            "partition": partition or ("eigenvector shards" if world > 1 else "single GPU")}


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port (oracle/mugiq_oracle.cpp) on the host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_loop_rate(wl, nev_sample, reps=1):
    """Times the oracle's restatement of the whole path on `nev_sample` eigenvectors of the workload (same lattice, same
    displacement entries, same momenta), all host threads: Loop_Mugiq::computeCoarseLoop in the reference's own schedule
    (stages 1+2), then convertIdxOrder_mapGamma and the projection GEMM (stages 3+4).
    Returns (contractions/s, threads, seconds)."""
    from oracle import oracle as orc
    from mugiq_b200 import synth
    from mugiq_b200.params import momenta_up_to
    L = wl["L"]
    entries = entries_of(wl)
    mom = momenta_up_to(wl["p2max"])
    ev = synth.random_evecs_np(L, nev_sample, seed=5)
    U = synth.random_gauge(L, seed=5) if entries else None
    sig = synth.sigmas(nev_sample)
    V3 = int(L[0]) * int(L[1]) * int(L[2])
    nLoop = loop_count(entries)
    ph = orc.phase_matrix(mom, -1, L)

    def one_pass(e, s, ent):
        pos = orc.compute_loop(e, s, U, ent, L)
        nl = loop_count(ent)
        mp = orc.reorder_mapgamma(pos, nl, L)
        return orc.gemm(mp, ph, int(L[3]) * 16 * nl, len(mom), V3)

    one_pass(ev[:1], sig[:1], entries[:1])  # warm-up (page faults, thread pool)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        one_pass(ev, sig, entries)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    V4 = int(np.prod(L))
    return nev_sample * V4 * nLoop / best, orc.num_threads(), best


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
        L = wl["L"]
        entries = entries_of(wl)
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
        out = {"value": nev_sample * V4 * loop_count(entries) / dt, "unit": UNIT, "kind": "reference kernels on this GPU",
               "seconds": dt,
               "sample": f"{nev_sample} of {wl['nev']} eigenvectors, full lattice, all {loop_count(entries)} loops; "
                         "lib/contract_wrappers.cu + lib/mugiq_*_kernels.cu unmodified (oracle/_ref), loop nest of "
                         "lib/loop_mugiq.cpp:455-509 restated around them, device printf of the contraction kernel sent to /dev/null"}
        ko = getattr(ref, "kernel_only_ms", None)
        if ko is not None:
            # the wall clock above is dominated by the wrappers' per-call cudaMalloc / cudaMemcpy / cudaDeviceSynchronize /
            # cudaFree and the in-kernel printf; this is the time of the reference's KERNELS alone (CUDA events around
            # the same launches, argument structs pre-staged)
            try:
                k = ko(evq, sig, gauge, entries, L)
                out["kernel_only"] = {"value": nev_sample * V4 * loop_count(entries) / (k["ms"] * 1e-3), "ms": k["ms"],
                                      "launches": k["launches"], "note": k["note"]}
            except Exception as exc:
                out["kernel_only"] = {"unavailable": repr(exc)[:200]}
        return out
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
    nloop = loop_count(entries_of(wl))
    mean_dt = float(np.mean(times))
    value = nev_s * V4 * nloop / mean_dt
    sample = (f"{nev_s} of {args.nev or wl['nev']} eigenvectors per step, full lattice, all loops, all four stages "
              "(oracle port, OpenMP)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": mean_dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(name, wl, args.nev or wl["nev"], args.gpus, args.evec_batch),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "reference_gpu": reference_gpu_rate(wl, min(wl["nev"], 20))}
    OUT.emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------------
class Harness:
    """Process-wide state of one bench run: device, ranks, barrier, the timed() primitive."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        from mugiq_b200.dist import env_rank_world
        self.rank, self.world, self.local_rank = env_rank_world()
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
        self.affinity0 = os.sched_getaffinity(0)
        self.numa_cores = bind_near_gpu(physical_device_index(self.local_rank))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.group = None
        self.comm = None
        if self.world > 1:
            # NCCL_DEBUG=VERSION makes NCCL print its version on stdout, which must carry exactly one JSON line
            if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"
            dist.init_process_group("nccl", device_id=self.dev)
            self.group = dist.group.WORLD

    def library_comm(self):
        """The library's own NCCL communicator over the ranks of the run (mugiq_b200_comm_*), made on first use."""
        if self.comm is None and self.world > 1:
            from mugiq_b200 import ops
            self.comm = ops.Comm(self.group, device=self.dev)
        return self.comm

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup, sampler=None):
        """W warm-up calls, then exactly `steps` calls between CUDA events, barrier + synchronize on both sides, max over
        ranks; the library's per-kernel event timers run over the timed region."""
        from mugiq_b200 import ops
        torch, dist = self.torch, self.dist
        for _ in range(warmup):
            fn()
        self.barrier()
        ops.prof_reset()
        ops.prof_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler is not None:
            sampler.__enter__()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        if sampler is not None:
            sampler.__exit__()
        ops.prof_enable(False)
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    def close(self):
        if self.comm is not None:
            self.torch.cuda.synchronize()
            self.comm.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def make_params(wl, U):
    from mugiq_b200.params import MugiqLoopParam, momenta_up_to
    prm = MugiqLoopParam(gauge=[U[mu] for mu in range(4)] if U is not None else None)
    if wl["entries"]:
        prm.set_displacements(wl["entries"])
    mom = momenta_up_to(wl["p2max"])
    prm.set_momenta(mom)
    return prm, mom


def pcie_ceiling(h):
    """Aggregate pinned host -> device bandwidth with ALL ranks copying at once (1 GiB each, best of 3): the ceiling of
    the end-to-end leg, whose step moves the whole eigenvector set over PCIe."""
    torch, dist = h.torch, h.dist
    n = 1 << 30
    src = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    src.zero_()
    dst = torch.empty(n, dtype=torch.uint8, device=h.dev)
    best = None
    for _ in range(4):
        h.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if h.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=h.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        best = ms if best is None else min(best, ms)
    del src, dst
    return h.world * n / (best * 1e-3) / 1e9


def verify_subset(h, wl, prm, mom, ev_d, sig, U, nsub=4):
    """The first `nsub` eigenvectors of the timed set through the same public path (Loop_Mugiq, default batching), checked
    against the CPU oracle: dataPos and dataMom to 1e-12 (norm-relative).  Rank-local (no cross-rank sum)."""
    from mugiq_b200.loop import Loop_Mugiq, Eigsolve
    from oracle import oracle as orc
    L = wl["L"]
    V3 = int(L[0]) * int(L[1]) * int(L[2])
    sub = Loop_Mugiq(prm, Eigsolve(list(ev_d[:nsub]), sig[:nsub], L), device=h.dev, evec_batch=h.args.evec_batch,
                     copy_pos_to_host=False)
    sub.computeCoarseLoop()
    entries = sub.cPrm.entries() if sub.cPrm.doNonLocal else []
    ev_np = ev_d[:nsub].cpu().numpy()
    ref = orc.compute_loop(ev_np, sig[:nsub], U if entries else None, entries, L)
    scale = np.abs(ref).max()
    err_pos = float(np.abs(sub.dataPos_d.cpu().numpy() - ref).max() / scale)
    nl = ref.shape[0]
    ref_mom = orc.gemm(orc.reorder_mapgamma(ref, nl, L), orc.phase_matrix(mom, -1, L), int(L[3]) * 16 * nl, len(mom), V3)
    ref_mom = ref_mom.reshape(len(mom), 16 * nl, int(L[3]))
    err_mom = float(np.abs(sub.dataMom.numpy() - ref_mom).max() / np.abs(ref_mom).max())
    return {"ok": bool(err_pos < 1e-12 and err_mom < 1e-12), "dataPos_rel_err": err_pos, "dataMom_rel_err": err_mom,
            "eigenvectors": nsub, "against": "oracle port (CPU), tolerance 1e-12"}


def trace_checksum(loop, sig_total):
    """Checksum of the buffer that was just timed: sum_x T_1(x) = sum_n |v_n|^2 / sigma_n = sum_n 1/sigma_n for the
    ultra-local identity loop (normalised eigenvectors), over every rank's shard."""
    got = complex(loop.dataPos_interior()[0, 0].sum().item()) if loop.tsplit is not None else complex(loop.dataPos_d[0, 0].sum().item())
    return {"sum_x_T1": got.real, "expected_sum_inv_sigma": sig_total, "rel_err": abs(got - sig_total) / abs(sig_total)}


def roofline_of(report, ms_step, steps, nmom):
    peaks, peak_src = measured_peaks()
    fpk = fp64_peaks()
    dom = max((k for k in report if k != "allreduce"), key=lambda k: report[k]["ms"], default=None)
    if dom is None or report[dom]["timed"] == 0:
        return None
    d = report[dom]
    per_launch_ms = d["ms"] / d["timed"]
    per_launch_bytes = d["alg_bytes"] / d["launches"]
    gbs = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get(dom, {}).get("dram_bytes_per_launch")
    hbm = {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "peak_source": peak_src}
    roof = dict(hbm, bound="hbm", kernel=dom)
    if d.get("alg_flops", 0) > 0:
        tf = d["alg_flops"] / d["launches"] / (per_launch_ms * 1e-3) / 1e12
        ridge = fpk["dfma_tflops"] * 1e3 / peaks["hbm_gbs"]  # flop per byte where the two roofs meet
        intensity = d["alg_flops"] / d["alg_bytes"]
        fp64 = {"achieved": tf, "peak": fpk["dfma_tflops"], "unit": "TFLOP/s", "frac": tf / fpk["dfma_tflops"],
                "peak_source": fpk["source"] + "; the driver's MEASURED_PEAKS.json holds no FP64 figure",
                "frac_of_dmma_stream": tf / fpk["dmma_tflops"], "frac_of_cublas_zgemm": tf / fpk["zgemm_4096_tflops"],
                "flop_per_algorithmic_byte": intensity, "ridge_flop_per_byte": ridge}
        if fpk.get("dfma_cplx_outer_tflops"):
            # the kernel's own instruction pattern as a bare register stream (tools/dfma_bench.cu, same 8 warps per SM): what
            # the FP64 pipe delivers for complex outer products, 2.52 cycles per DFMA against 2.22 for a chained stream
            fp64["frac_of_cplx_outer_product_stream"] = tf / fpk["dfma_cplx_outer_tflops"]
            fp64["cplx_outer_product_stream"] = {"tflops": fpk["dfma_cplx_outer_tflops"],
                                                 "cycles_per_dfma": fpk.get("dfma_cplx_outer_cycles"),
                                                 "chained_stream_cycles_per_dfma": fpk.get("dfma_chain_cycles"),
                                                 "source": "tools/dfma_bench.cu (profiles/r3_dfma_bench.json)"}
        if intensity > ridge:
            # FP64-pipe bound: the headline roofline is the FP64 one; the HBM figures stay beside it
            roof = {"bound": "fp64", "kernel": dom, "achieved": tf, "peak": fpk["dfma_tflops"], "unit": "TFLOP/s",
                    "frac": tf / fpk["dfma_tflops"], "peak_source": fp64["peak_source"], "fp64": fp64, "hbm": hbm}
        else:
            roof["fp64"] = fp64
    roof.update({"traffic": traffic, "alg_bytes_per_launch": per_launch_bytes, "ms_per_launch": per_launch_ms,
                 "share_of_step": d["ms"] / (ms_step * steps),
                 "kernels": {k: {"launches": v["launches"], "ms": round(v["ms"], 4),
                                 "GBps": (v["alg_bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 else None}
                             for k, v in report.items()}})
    # stage 4 (momentum projection) separately, as SURVEY 8d asks: useful FP64 TFLOP/s (8*M*N*K) against the measured
    # DMMA peak; with few momenta (N <~ 12) the projection is HBM bound and the GB/s figure above is the relevant one
    mp = report.get("momproj")
    if mp and mp["ms"] > 0 and mp.get("alg_flops", 0) > 0:
        tf = mp["alg_flops"] / (mp["ms"] * 1e-3) / 1e12
        roof["projection"] = {"TFLOPs": tf, "dmma_peak": fpk["dmma_tflops"], "frac": tf / fpk["dmma_tflops"], "Nmom": nmom,
                              "bound": "tensor" if nmom >= 12 else "hbm",
                              "GBps": mp["alg_bytes"] / (mp["ms"] * 1e-3) / 1e9, "hbm_frac": mp["alg_bytes"] / (mp["ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                              "peak_source": fpk["source"]}
    return roof


def setup_tsplit(h, wl, nev, halo):
    """This rank's time slab of wl['L'] out of a global lattice with T = L[3] * world: TSplit, device links of the extended
    slab, eigenvector slabs in the extended layout (peer-mapped for the NVLink transports).  Returns (ts, U, es, cleanup)."""
    torch, dist = h.torch, h.dist
    from mugiq_b200 import ops, synth
    from mugiq_b200.loop import Eigsolve
    from mugiq_b200.tsplit import TSplit
    from mugiq_b200.params import parse_disp_entries
    L = wl["L"]
    tmax = 0
    if wl["entries"]:
        _, ds, _, stop = parse_disp_entries(wl["entries"])
        tmax = max([b for s, b in zip(ds, stop) if s[1] == "t"] + [0])
    Lg = (L[0], L[1], L[2], L[3] * h.world)
    ts = TSplit(Lg, h.rank, h.world, tmax)
    # this rank's extended slab of one global field, generated on the device slice by slice (slice-keyed seeds)
    U = synth.random_gauge_slab_torch(Lg, [(ts.t0 - ts.H + i) % Lg[3] for i in range(ts.Tl + 2 * ts.H)], seed=11, device=h.dev)
    sig = synth.sigmas(nev)
    opened = []
    if halo == "nccl":
        ev_d = synth.random_evecs_torch(ts.L_ext, nev, seed=100 + h.rank, device=h.dev)
    else:
        # NVLink peer mode: the slabs live in one IPC-shareable allocation the two time neighbours map; halos are
        # written straight into the neighbours' slabs (copy engines or SM kernel)
        V4e = int(np.prod(ts.L_ext))
        peer_buf = ops.PeerBuffer(nev * V4e * 12 * 16, device=h.dev)
        ev_d = synth.random_evecs_torch(ts.L_ext, nev, seed=100 + h.rank, device=h.dev,
                                        out=peer_buf.tensor((nev, V4e, 12), torch.complex128))
        if h.world > 1:
            handles = [None] * h.world
            dist.all_gather_object(handles, peer_buf.handle)
            up_ptr = ops.peer_open(handles[(h.rank + 1) % h.world], h.dev)
            dn_ptr = up_ptr if h.world == 2 else ops.peer_open(handles[(h.rank - 1) % h.world], h.dev)
            opened = list({up_ptr, dn_ptr})
        else:
            up_ptr = dn_ptr = peer_buf.ptr
        ts.attach_peers(ev_d, up_ptr, dn_ptr, mode=0 if halo == "dma" else 1, group=h.group)
    es = Eigsolve(list(ev_d), sig, L, ext_volume=int(np.prod(ts.L_ext)))

    def cleanup():  # unmap the neighbours' slabs before anybody frees its own
        torch.cuda.synchronize()
        if h.world > 1 and opened:
            dist.barrier()
            for ptr in opened:
                ops.peer_close(ptr, h.dev)
            dist.barrier()

    return ts, U, es, sig, cleanup


def leg_config4(h, steps, warmup):
    """BASELINE configs[3]: 32^3x64, 1000 eigenvectors over 8 GPUs = 125 per GPU (50.3 GB), ultra-local + the 8 one-hop
    loops, and the POSITION-SPACE loop buffer (5 computed loops = 2.7 GB per GPU) summed over the GPUs: once with the
    library's chunked all-reduce overlapping the kernels (loop_plan_accumulate_allreduce), once as one all-reduce after the
    kernels, once without (only the projected 0.5 MB buffer is summed)."""
    torch = h.torch
    from mugiq_b200 import ops, synth
    from mugiq_b200.loop import Loop_Mugiq, Eigsolve
    name = "32x32x32x64_nev125_ulocal+1hop8"
    wl = WORKLOADS[name]
    L, nev = wl["L"], (h.args.config4_nev or wl["nev"])
    V4 = int(np.prod(L))
    U = synth.random_gauge_slab_torch(L, list(range(L[3])), seed=13, device=h.dev)   # replicated links, made on the device
    prm, mom = make_params(wl, U)
    sig = synth.sigmas(nev) + 0.2 * h.rank
    ev_d = synth.random_evecs_torch(L, nev, seed=300 + h.rank, device=h.dev)
    es = Eigsolve(list(ev_d), sig, L)
    out = {"workload": name, "L": list(L), "nev_per_gpu": nev, "nLoop": 9, "Nmom": len(mom), "n_gpus": h.world, "steps": steps}
    units = h.world * nev * V4 * 9
    variants = [("without_pos_allreduce", dict(reduce_pos=False))]
    if h.world > 1:
        variants += [("pos_allreduce_overlapped", dict(reduce_pos=True, comm=h.library_comm(), allreduce_chunks=h.args.allreduce_chunks)),
                     ("pos_allreduce_after_kernels", dict(reduce_pos=True, comm=h.library_comm(), allreduce_chunks=1)),
                     # the same overlapped sum with the chunks moved by the copy engines over peer-mapped buffers
                     ("pos_allreduce_peer_dma", dict(reduce_pos=True, comm=h.library_comm(), allreduce_chunks=h.args.allreduce_chunks,
                                                     peer_reduce=True))]
    for label, kw in variants:
        try:
            loop = Loop_Mugiq(prm, es, device=h.dev, group=h.group, evec_batch=256, copy_pos_to_host=False, **kw)
            loop.MomProjDone = False
            loop.computeCoarseLoop()   # first step outside the timed region (plan, peer mapping)
        except Exception as exc:       # a transport this box cannot do (no IPC between the devices) must not cost the other forms
            out[label] = {"error": repr(exc)[:300]}
            continue

        def step():
            loop.MomProjDone = False
            loop.computeCoarseLoop()

        ms = h.timed(step, steps, warmup)
        rep = ops.prof_report()
        o = {"ms_per_step": ms, "value": units / (ms * 1e-3), "loop_fused_ms_per_step": rep.get("loop_fused", {}).get("ms", 0.0) / steps}
        ar = rep.get("allreduce")
        if ar and ar["ms"] > 0 and label != "without_pos_allreduce":
            nbytes = ar["alg_bytes"] / steps
            o["allreduce_ms_per_step"] = ar["ms"] / steps
            o["allreduce_bytes_per_step"] = nbytes
            o["allreduce_launches_per_step"] = ar["launches"] / steps
            busbw = nbytes * 2 * (h.world - 1) / h.world / (ar["ms"] / steps * 1e-3) / 1e9
            o["allreduce_bus_GBps"] = busbw
            o["allreduce_bus_frac_of_725"] = busbw / 725.0
        if label == "without_pos_allreduce":
            o["note"] = "position-space buffer stays rank-local, the projected buffer (16*Nmom*T*nLoop complex) is summed"
        else:
            s_tot = float((1.0 / sig).sum())
            t = torch.tensor([s_tot], dtype=torch.float64, device=h.dev)
            h.dist.all_reduce(t)
            got = complex(loop.dataPos_d[0, 0].sum().item())
            o["checksum_rel_err"] = abs(got - float(t.item())) / float(t.item())   # summed buffer: sum_x T_1 = sum over ALL ranks of 1/sigma
        out[label] = o
        loop.close_peer_reduce()
        del loop
    best = out["without_pos_allreduce"]
    summed = [(k, out[k]) for k in ("pos_allreduce_overlapped", "pos_allreduce_peer_dma")
              if "ms_per_step" in out.get(k, {}) and out[k].get("checksum_rel_err", 1.0) < 1e-10]
    if summed:  # the faster of the two overlapped forms whose summed buffer checked out
        k, best = min(summed, key=lambda kv: kv[1]["ms_per_step"])
        out["transport"] = ("copy engines over peer-mapped buffers (mugiq_b200_comm_attach_peers)" if k == "pos_allreduce_peer_dma"
                            else "NCCL all-reduce kernels")
    out["value"] = best["value"]
    out["ms_per_step"] = best["ms_per_step"]
    out["unit"] = UNIT
    out["note"] = ("value = the step WITH the position-space all-reduce (the faster overlapped form) at N > 1; weak scaling, so "
                   "efficiency = value(N) / (N * value(1))")
    del ev_d, es, U
    torch.cuda.empty_cache()
    return out


def leg_tsplit(h, steps, warmup):
    """BASELINE configs[4] at N > 1: the 48^3 x (12*N) lattice split in T, one 48^3x12 slab per GPU (+2 halo slices on each
    side), ultra-local + 8 one-hop loops, halos written into the neighbours' slabs over NVLink by the copy engines."""
    torch = h.torch
    from mugiq_b200 import ops
    from mugiq_b200.loop import Loop_Mugiq
    name = "48x48x48x12_nev300_ulocal+1hop8"
    wl = WORKLOADS[name]
    nev = h.args.tsplit_nev
    ts, U, es, sig, cleanup = setup_tsplit(h, wl, nev, "dma")
    prm, mom = make_params(wl, U)
    loop = Loop_Mugiq(prm, es, device=h.dev, group=h.group, evec_batch=256, copy_pos_to_host=False, tsplit=ts,
                      stream_batch=h.args.tsplit_batch)
    V4 = int(np.prod(wl["L"]))

    def step():
        loop.MomProjDone = False
        loop.computeCoarseLoop()

    ms = h.timed(step, steps, warmup)
    rep = ops.prof_report()
    out = {"workload": name, "L_slab": list(wl["L"]), "L_global": list(ts.L_global), "nev": nev, "nLoop": loop.cPrm.nLoop,
           "n_gpus": h.world, "halo": "dma", "ms_per_step": ms, "value": h.world * nev * V4 * loop.cPrm.nLoop / (ms * 1e-3), "unit": UNIT,
           "loop_fused_ms_per_step": rep.get("loop_fused", {}).get("ms", 0.0) / steps}
    hp = rep.get("halo_push")
    if hp and hp["ms"] > 0:
        # the library counts a push as read + write; what crosses NVLink is half of that, one way (upper halos only:
        # the minus-t loop is derived from its plus partner)
        out["halo_bytes_sent_per_step"] = hp["alg_bytes"] / 2 / steps
        out["halo_push_ms_per_step"] = hp["ms"] / steps
        out["halo_GBps_per_direction"] = hp["alg_bytes"] / 2 / (hp["ms"] * 1e-3) / 1e9
        out["halo_frac_of_770"] = out["halo_GBps_per_direction"] / 770.0
    # configs[4] names 2000 eigenvectors: 510 GB of slabs per GPU, more than HBM holds, so in production a producer (QUDA's
    # prolongator through setEvecProducer) refills the slab batches.  Here the resident slabs are recycled: the step below
    # pushes all 2000 eigenvector slabs (10 passes over the resident 200, each with its own sigma) through halo pushes and
    # kernels - the full-size step without the producer's own time.
    passes = max(1, wl_full_nev() // nev)
    if passes > 1 and not h.args.no_tsplit_full:
        try:
            from mugiq_b200.loop import Eigsolve
            from mugiq_b200 import synth
            es_full = Eigsolve(list(es.eVecs) * passes, synth.sigmas(nev * passes), es.L, ext_volume=es.ext_volume)
            loop_f = Loop_Mugiq(prm, es_full, device=h.dev, group=h.group, evec_batch=256, copy_pos_to_host=False, tsplit=ts,
                                stream_batch=h.args.tsplit_batch)

            def step_full():
                loop_f.MomProjDone = False
                loop_f.computeCoarseLoop()

            ms_f = h.timed(step_full, 2, 1)
            out["full_eigenvector_count"] = {"nev": nev * passes, "passes_over_resident_slabs": passes, "ms_per_step": ms_f,
                                             "value": h.world * nev * passes * V4 * loop.cPrm.nLoop / (ms_f * 1e-3), "unit": UNIT,
                                             "note": "all eigenvector slabs of configs[4] through halo pushes and kernels in one "
                                                     "step; the resident slabs are recycled (a producer would refill them)"}
            del loop_f, es_full
        except Exception as exc:
            out["full_eigenvector_count"] = {"error": repr(exc)[:300]}
    del loop
    cleanup()
    del es, U
    torch.cuda.empty_cache()
    return out


def wl_full_nev():
    """eigenvector count BASELINE.json configs[4] names ("48^3x96 lattice, 2000 eigvecs, T-split ... on 8xB200")"""
    return 2000


def leg_e2e_cpp(h, wl, name):
    """The same end-to-end step through the C++ front end (mugiq_b200/host: Loop_Mugiq<double> + the streamed
    eigenvector feed over the C-ABI), run as its own process: loop_driver --bench prints the e2e fields."""
    exe = os.path.join(ROOT, "mugiq_b200", "host", "build", "loop_driver")
    if not os.path.exists(exe):
        return {"unavailable": "mugiq_b200/host/build/loop_driver not built"}
    L = wl["L"]
    cmd = [exe, "--bench", "--dim", *[str(x) for x in L], "--n-ev", str(h.args.nev or wl["nev"]), "--bench-steps", "3",
           "--bench-batch", str(h.args.stream_batch), "--device", str(h.local_rank)]
    if wl["entries"]:
        cmd += ["--displace-entry-string", wl["entries"]]
    cmd += ["--bench-p2max", str(wl["p2max"])]
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        for ln in res.stdout.splitlines()[::-1]:
            if ln.startswith("{"):
                return json.loads(ln)
        return {"unavailable": ("rc=%d " % res.returncode) + (res.stderr or res.stdout)[-300:]}
    except Exception as exc:
        return {"unavailable": repr(exc)[:300]}


def run_ours(args, wl, name):
    h = Harness(args)
    torch, dist = h.torch, h.dist
    from mugiq_b200 import ops, synth
    from mugiq_b200.loop import Loop_Mugiq, Eigsolve

    rank, world, dev, group = h.rank, h.world, h.dev, h.group
    L, nev = wl["L"], args.nev or wl["nev"]
    V4 = int(np.prod(L))
    ts = None
    cleanup = None
    if args.tsplit:
        # lattice-T split: every rank owns a slab of wl["L"] of a global lattice with T = L[3] * world; the eigenvectors
        # span all ranks (same sigma everywhere), the halos of +-t displacements travel over NVLink
        ts, U, es, sig, cleanup = setup_tsplit(h, wl, nev, args.halo)
        ev_d = None
    else:
        U = synth.random_gauge(L, seed=11)
        sig = synth.sigmas(nev) + 0.2 * rank
        ev_d = synth.random_evecs_torch(L, nev, seed=100 + rank, device=dev)      # [nev, V4, 12] resident in HBM
        es = Eigsolve(list(ev_d), sig, L)
    prm, mom = make_params(wl, U)
    loop = Loop_Mugiq(prm, es, device=dev, group=group, evec_batch=args.evec_batch,
                      copy_pos_to_host=False, tsplit=ts, stream_batch=args.tsplit_batch if ts is not None else 16)
    nLoop = loop.cPrm.nLoop
    units_per_rank = nev * V4 * nLoop

    # ---- leg 1: inputs resident in HBM --------------------------------------------------------------------
    def step_resident():
        loop.MomProjDone = False
        loop.computeCoarseLoop()

    sampler = ClockSampler(physical_device_index(h.local_rank))
    ms_step = h.timed(step_resident, args.steps, args.warmup, sampler)
    halo_sides = getattr(loop, "tsplit_halo_sides", 2)
    halo_slices = getattr(loop, "tsplit_halo_slices", 2 * ts.H if ts is not None else 0)
    if ts is not None and getattr(loop, "_trace_on", False):
        torch.cuda.synchronize()
        print(f"[rank {rank}] T-split phases of the last step (device ms, host ms): " +
              "; ".join(f"{n} {g:.3f}/{hh:.3f}" for n, g, hh in loop.trace_report()), file=sys.stderr, flush=True)
    report = ops.prof_report()
    launches = sum(v["launches"] for k, v in report.items() if k != "allreduce")  # NCCL's kernels are not ours
    value = world * units_per_rank / (ms_step * 1e-3)
    roofline = roofline_of(report, ms_step, args.steps, len(mom))

    # ---- the buffer that was just timed: checksum, and a 4-eigenvector subset against the oracle ------------
    verified = None
    if not args.no_verify:
        try:
            s_tot = float((1.0 / sig).sum())
            if ts is not None:   # every rank's slab holds Tl of the T slices of the same eigenvectors, normalised per slab
                chk = None
            else:
                # the projected buffer is what the ranks summed; the position-space buffer is rank-local
                chk = trace_checksum(loop, s_tot)
            if rank == 0 and ts is None:
                verified = verify_subset(h, wl, prm, mom, ev_d, sig, U)
                verified["timed_buffer_checksum"] = chk
                verified["ok"] = bool(verified["ok"] and chk["rel_err"] < 1e-10)
        except Exception as exc:
            verified = {"ok": False, "error": repr(exc)[:300]}

    # ---- leg 2: end to end through the public API with HOST buffers -----------------------------------------
    e2e = None
    if not args.no_e2e and ts is None:
        ev_h = torch.empty((nev, V4, 12), dtype=torch.complex128, pin_memory=True)
        ev_h.copy_(ev_d)
        del loop
        # the links too live in PINNED host memory (numpy views of one pinned tensor): from pageable memory their 75 MB cost
        # 5 ms of staged copy per step (profiles/r3_pcie_big.json: the eigenvector + dataPos copies alone need 97.7 ms)
        prm_h = prm
        if U is not None:
            U_pin = torch.from_numpy(np.ascontiguousarray(U)).pin_memory()
            prm_h, _ = make_params(wl, U_pin.numpy())
        loop_h = Loop_Mugiq(prm_h, Eigsolve(list(ev_h), sig, L), device=dev, group=group, evec_batch=args.evec_batch,
                            copy_pos_to_host=True, stream_batch=args.stream_batch)

        def step_host():
            loop_h.MomProjDone = False
            loop_h.displace.upload_gauge(prm_h) if loop_h.displace is not None else None   # H2D of the gauge field
            loop_h.computeCoarseLoop()                                                    # H2D evecs, D2H dataPos + dataMom

        e2e_steps = max(1, min(args.steps, 5))
        ms_e2e = h.timed(step_host, e2e_steps, 1)
        h2d = ev_h.numel() * 16 + (U.nbytes if loop_h.displace is not None else 0)
        d2h = loop_h.dataPos.numel() * 16 + loop_h.dataMom.numel() * 16
        e2e = {"value": world * units_per_rank / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e, "steps": e2e_steps,
               "h2d_GBps_per_gpu": h2d / (ms_e2e * 1e-3) / 1e9, "cpu_cores_bound_near_gpu": h.numa_cores}
        del loop_h
        try:
            ceil = pcie_ceiling(h)
            e2e["pcie_ceiling_gbs"] = ceil
            e2e["pcie_ceiling_note"] = (f"aggregate pinned H2D with all {world} rank(s) copying 1 GiB at once; the step's H2D alone "
                                        f"needs {world * h2d / (ceil * 1e9) * 1e3:.1f} ms at that rate")
            e2e["frac_of_pcie_ceiling"] = (world * h2d / (ms_e2e * 1e-3) / 1e9) / ceil
        except Exception as exc:
            e2e["pcie_ceiling_gbs"] = None
            e2e["pcie_ceiling_note"] = repr(exc)[:200]

    # ---- leg 3: eigenvectors resident in QUDA's native FLOAT2 order (what a QUDA-backed caller holds) ----------------
    # the fused kernel stages the FLOAT2 fields itself (tensor boxes): no conversion pass, no site-major copy
    quda = None
    if not args.no_e2e and ts is None and world == 1:
        ev_q = torch.empty_like(ev_d)
        for i in range(nev):
            ev_q[i] = ops.export_spinor(ev_d[i], 2, L)
        loop_q = Loop_Mugiq(prm, Eigsolve(list(ev_q), sig, L, field_order=2), device=dev, group=group, evec_batch=args.evec_batch,
                            copy_pos_to_host=False)

        def step_quda():
            loop_q.MomProjDone = False
            loop_q.computeCoarseLoop()

        q_steps = max(1, min(args.steps, 5))
        ms_q = h.timed(step_quda, q_steps, 2)
        rep_q = ops.prof_report()
        quda = {"value": world * units_per_rank / (ms_q * 1e-3), "unit": UNIT, "ms_per_step": ms_q, "steps": q_steps,
                "loop_fused_ms_per_step": rep_q.get("loop_fused", {}).get("ms", 0.0) / q_steps,
                "checksum": trace_checksum(loop_q, float((1.0 / sig).sum())),
                "note": "eigenvectors resident in HBM in QUDA FLOAT2 order; the fused kernel stages them itself as TMA tensor "
                        "boxes (mugiq_b200_loop_plan_set_evec_order): no layout conversion inside the step"}
        del loop_q, ev_q

    # free the main workload before the extra legs
    if ts is not None:
        del loop
        cleanup()
    ev_d = es = None
    torch.cuda.empty_cache()

    # ---- BASELINE configs[3] and configs[4] beside the default workload ------------------------------------------------
    config4 = tsplit_obj = None
    if not args.no_extra and ts is None and name == DEFAULT_WORKLOAD:
        try:
            config4 = leg_config4(h, max(2, min(args.steps, 4)), 2)
        except Exception as exc:
            config4 = {"error": repr(exc)[:300]}
            torch.cuda.empty_cache()
        if world > 1 and not args.no_tsplit_extra:
            try:
                tsplit_obj = leg_tsplit(h, max(2, min(args.steps, 4)), 2)
            except Exception as exc:
                tsplit_obj = {"error": repr(exc)[:300]}
                torch.cuda.empty_cache()

    # ---- cpu baseline and the C++ front end (rank 0, N == 1 only) ---------------------------------------------------
    cpu = ref_gpu = e2e_cpp = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as orc
        os.sched_setaffinity(0, h.affinity0)  # the CPU baseline uses every host core again
        orc.set_num_threads(len(os.sched_getaffinity(0)))
        nev_s = cpu_sample_size(wl, 12.0)
        rate, threads, dt = cpu_loop_rate(wl, nev_s)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{nev_s} of {nev} eigenvectors, full lattice, all {nLoop} loops, all four stages, {dt:.1f} s (oracle port, OpenMP)"}
        ref_gpu = reference_gpu_rate(wl, min(nev, 20))
    if rank == 0 and world == 1 and not args.no_e2e and ts is None:
        e2e_cpp = leg_e2e_cpp(h, wl, name)

    if rank == 0:
        partition = None
        if args.tsplit:
            partition = ("lattice-T split, global T = %d, halo %d slices, %.1f MB of halo per rank and step over "
                         "NVLink (%d-sided eigenvector halo of %d slice(s), transport %s; interior-only compute)"
                         % (L[3] * world, ts.H, nev * ts.halo_bytes_per_vector(slices=halo_slices) / 1e6, halo_sides,
                            halo_slices, args.halo))
        clk = sampler.summary()
        if roofline and roofline.get("bound") == "fp64" and clk.get("sm_mhz") and clk.get("sm_max_mhz"):
            # under the board's power cap a long FP64 step runs below the maximum SM clock: the same fraction per cycle
            roofline["fp64"]["frac_clock_normalised"] = roofline["fp64"]["frac"] * clk["sm_max_mhz"] / clk["sm_mhz"]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": make_config(name, wl, nev, world, args.evec_batch, partition),
                "roofline": roofline, "verified": verified["ok"] if verified else None, "verification": verified,
                "cpu_baseline": cpu, "reference_gpu": ref_gpu, "e2e": e2e, "e2e_cpp": e2e_cpp, "quda_order": quda,
                "config4": config4, "tsplit": tsplit_obj, "gpu_launches": launches, "clocks": clk}
        OUT.emit(line)
    h.close()
    return 0


class OneLineStdout:
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version from ncclCommInitRank
    at NCCL_DEBUG=VERSION and =WARN), so file descriptor 1 is pointed at stderr for the duration of the run and the line
    goes to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, obj):
        sys.stdout.flush()
        os.write(self.real, (json.dumps(obj) + "\n").encode())


OUT = None


def main():
    global OUT
    OUT = OneLineStdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS))
    ap.add_argument("--nev", type=int, default=0, help="override the eigenvector count per GPU (debugging)")
    ap.add_argument("--evec-batch", type=int, default=200)
    ap.add_argument("--stream-batch", type=int, default=16, help="eigenvectors per H2D staging buffer of the end-to-end leg")
    ap.add_argument("--tsplit", action="store_true", help="partition the lattice in T over the GPUs (halo exchange) instead of "
                                                          "sharding eigenvectors")
    ap.add_argument("--halo", default="dma", choices=["dma", "kernel", "nccl"],
                    help="--tsplit halo transport: direct NVLink writes into the neighbours' slabs by the copy engines (dma) or "
                         "an SM push kernel (kernel), or NCCL send/recv through staging buffers (nccl)")
    ap.add_argument("--tsplit-batch", type=int, default=100, help="--tsplit: eigenvectors per halo push / kernel launch (the push "
                                                                  "of batch i+1 overlaps the kernels of batch i)")
    ap.add_argument("--tsplit-nev", type=int, default=200, help="eigenvectors per GPU of the `tsplit` object (68 GB of slabs)")
    ap.add_argument("--config4-nev", type=int, default=0, help="eigenvectors per GPU of the `config4` object (default 125)")
    ap.add_argument("--allreduce-chunks", type=int, default=8, help="time-slice chunks of the overlapped position-space all-reduce")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the config4 / tsplit objects")
    ap.add_argument("--no-tsplit-extra", action="store_true", help="skip the `tsplit` object (keep `config4`)")
    ap.add_argument("--no-tsplit-full", action="store_true", help="skip the 2000-eigenvector pass of the `tsplit` object")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl, args.workload)
    return run_ours(args, wl, args.workload)


if __name__ == "__main__":
    sys.exit(main())
