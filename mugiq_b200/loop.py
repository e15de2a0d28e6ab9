"""Python mirror of the reference's operator interface for the hot path: `Loop_Mugiq`
(/root/reference/include/loop_mugiq.h:12-136, lib/loop_mugiq.cpp) and `Displace`
(include/displace.h:13-100, lib/displace.cpp).  Method names, argument meaning, buffer index orders and
error behaviour follow the reference; all arithmetic happens in the CUDA library behind the C-ABI.

What differs on purpose:
  * the eigenvector x displacement loop nest is one C-ABI call per eigenvector batch instead of
    4-6 synchronous launches and 3 field copies per eigenvector and hop;
  * eigenvectors may be sharded over ranks (data-parallel over n); the loop buffer is then summed with one
    NCCL allreduce instead of MPI_Reduce/Gather/Bcast on host buffers (lib/loop_mugiq.cpp:406-424);
  * QUDA's eigensolver is an external input: `Eigsolve` only carries eVecs / eVals_sigma.
"""
import warnings
from typing import List, Optional

import numpy as np
import torch

from . import ops
from .dist import allreduce_loop_buffer
from ._lib import PREC_DOUBLE, PREC_SINGLE
from .lattice import Lattice
from .params import (MugiqLoopParam, LoopComputeParam, MugiqError, which_displace, DISPLACE_TYPE_COVARIANT,
                     GAMMA_NAMES)


class Eigsolve:
    """Stands in for Eigsolve_Mugiq as far as the loop path reads it (include/eigsolve_mugiq.h): the
    eigenvector fields `eVecs[n]` (device tensors [volume, 12] complex, canonical site-major even/odd order,
    or HOST tensors when the loop streams them) and `eVals_sigma[n]` (host floats, sigma_n = sqrt(lambda_n),
    lib/eigsolve_mugiq.cpp:289-315).  `L` are the local lattice extents the fields live on."""

    def __init__(self, eVecs, eVals_sigma, L, ext_volume=0, field_order=0):
        # ext_volume: sites of the T-split extended slab when the fields are stored with their halo slices allocated
        self.ext_volume = int(ext_volume)
        # field_order: 0 = canonical site-major, 2 = QUDA FLOAT2 ([parity][spin*3+colour][x_cb]; what
        # computeLoop dispatches on, lib/interface_mugiq.cpp:226-235): staged by the fused kernel itself, no conversion
        self.field_order = int(field_order)
        if self.field_order not in (0, 2):
            raise MugiqError("Eigsolve: field_order must be 0 (site-major) or 2 (QUDA FLOAT2); convert FLOAT4 fields with "
                             "ops.ingest_spinor_batch")
        self.eVecs = list(eVecs)
        self.eVals_sigma = [float(s) for s in eVals_sigma]
        self.L = tuple(int(x) for x in L)
        self.nEv = len(self.eVecs)
        if self.nEv == 0:
            raise MugiqError("Eigsolve: no eigenvectors")
        if len(self.eVals_sigma) != self.nEv:
            raise MugiqError("Eigsolve: eVals_sigma length does not match the number of eigenvectors")
        vol = Lattice(self.L).volume
        for v in self.eVecs:
            if (v.numel() != vol * 12 and v.numel() != self.ext_volume * 12) or not v.is_complex():
                raise MugiqError("Eigsolve: eigenvectors must be complex fields of volume*12 elements "
                                 "(full site subset, lib/contract_wrappers.cu:100)")

    @property
    def dtype(self):
        return self.eVecs[0].dtype


class Displace:
    """Displace<F,order> (include/displace.h:13-100).  Owns the device gauge field and the auxiliary
    displaced vector; the direction state machine is setupDisplacement -> doVectorDisplacement."""

    def __init__(self, loopParams: MugiqLoopParam, L, dtype=torch.complex128, device="cuda"):
        if loopParams.gauge is None:
            raise MugiqError("Displace: loopParams.gauge is not set")
        g0 = loopParams.gauge[0]
        want = dtype if isinstance(g0, torch.Tensor) else (np.complex128 if dtype == torch.complex128 else np.complex64)
        if not isinstance(g0, torch.Tensor):
            g0 = np.asarray(g0)
        # lib/displace.cpp:84-87: gauge precision must equal the template precision
        if g0.dtype != want:
            raise MugiqError(f"createCudaGaugeField: Incompatible precision settings between Displace template "
                             f"{dtype} and gauge field parameters {g0.dtype}")
        self.L = tuple(int(x) for x in L)
        self.device = device
        self.gaugeField = None
        self.gaugeVersion = 0
        self.upload_gauge(loopParams)
        self.auxDispVec = torch.zeros((Lattice(self.L).volume, 12), dtype=dtype, device=device)
        self.dispString = ""
        self.dispDir = None
        self.dispSign = None

    def upload_gauge(self, loopParams: MugiqLoopParam):
        """createCudaGaugeField (lib/displace.cpp:70-100): H2D copy of the host QDP-order links.  Links that already
        live on the device (four [V4, 3, 3] CUDA tensors, e.g. a time slab generated in place) are adopted."""
        if isinstance(loopParams.gauge[0], torch.Tensor) and loopParams.gauge[0].is_cuda:
            self.gaugeField = torch.stack([loopParams.gauge[mu] for mu in range(4)]).contiguous()
            self.gaugeVersion += 1
            return self.gaugeField
        if self.gaugeField is None:
            self.gaugeField = ops.gauge_upload(loopParams.gauge, self.L, device=self.device)
        else:
            ops.gauge_upload(loopParams.gauge, self.L, device=self.device, out=self.gaugeField)
        self.gaugeVersion += 1
        return self.gaugeField

    def setupDisplacement(self, dStr: str):
        self.dispString = dStr
        self.dispDir, self.dispSign = which_displace(dStr)  # raises like WhichDisplaceFlag

    def doVectorDisplacement(self, dispType, displacedEvec, idisp):
        """One hop in place: displacedEvec <- U * shift(displacedEvec) (lib/displace.cpp:55-67).  The
        reference's blas::zero + two field copies are replaced by a pointer swap of the storages."""
        if dispType != DISPLACE_TYPE_COVARIANT:
            raise MugiqError(f"Unsupported Displacement type {dispType}")
        if self.dispDir is None:
            raise MugiqError("doVectorDisplacement: setupDisplacement was not called")
        ops.displace(self.auxDispVec, displacedEvec, self.gaugeField, self.dispDir, self.dispSign, self.L)
        # swapAuxDispVec: exchange the underlying storages so the caller's tensor holds the result
        tmp = displacedEvec.data
        displacedEvec.data = self.auxDispVec.data
        self.auxDispVec.data = tmp
        return displacedEvec


class Loop_Mugiq:
    """Loop_Mugiq<Float,order> (include/loop_mugiq.h:12-136).

    Buffers (same index orders as the reference):
      dataPos_d  [nLoop, 16, V4]        x_eo + V4*(G + 16*iL)                   lib/loop_mugiq.cpp:468,492
      dataPosMP_d[V3, nData, Lt]        t + Lt*idata + Lt*nData*v3              lib/mugiq_util_kernels.cu:88-97
      dataMom_d  [Nmom, nData, Lt]      t + Lt*idata + Lt*nData*im              lib/loop_mugiq.cpp:415-418
    `dataPos`, `dataMom` are the host copies; `dataMom_bcast` equals dataMom for a single time block.
    `group` (optional torch.distributed process group): the eigenvectors given to this rank are its shard
    of the full set; loop buffers are summed over the group after the local eigenvector loop.
    """

    def __init__(self, loopParams_: MugiqLoopParam, eigsolve_: Eigsolve, device=None, group=None, evec_batch=64,
                 stream_batch=16, copy_pos_to_host=True, fused_momproj=True, tsplit=None, comm=None, reduce_pos=None,
                 allreduce_chunks=8, peer_reduce=False):
        self.eigsolve = eigsolve_
        # with `comm`: the overlapped position-space sum moves its chunks with the copy engines over peer-mapped buffers
        # (mugiq_b200_comm_attach_peers) instead of NCCL's all-reduce kernels; needs `group` for the handle exchange
        self.peer_reduce = bool(peer_reduce) and comm is not None and comm.size > 1
        self._peer = None
        self.group = group
        # ops.Comm over the same ranks as `group`: the cross-rank sums then go through the library's own NCCL calls
        # (mugiq_b200_allreduce*), the position-space one overlapped with the kernels chunk by chunk; without it they
        # are torch.distributed all-reduces after the kernels
        self.comm = comm
        # eigenvector shards: sum the position-space buffer over the ranks (True), or only the projected one (False);
        # None = the position-space buffer exactly when it is needed on every rank (host copy, or no projection)
        self.reduce_pos = reduce_pos
        self.allreduce_chunks = int(allreduce_chunks)
        self.evec_batch = int(evec_batch)      # eigenvectors per C-ABI call when they are device resident
        self.stream_batch = int(stream_batch)  # eigenvectors per H2D staging buffer when they live on the host
        # the reference always copies dataPos_d to the host (lib/loop_mugiq.cpp:512); a caller that only wants
        # momentum-space data can switch the 16*V4*nLoop-complex D2H copy off
        self.copy_pos_to_host = bool(copy_pos_to_host)
        # stages 3+4 as one kernel reading dataPos in place (mugiq_b200_momproj_pos) instead of the reference's
        # convertIdxOrder_mapGamma + GEMM pair; False keeps the two-call form (and its dataPosMP buffer)
        self.fused_momproj = bool(fused_momproj)
        # lattice-T split (mugiq_b200.tsplit.TSplit): the eigenvectors given are this rank's time-slab in local even/odd
        # order, `loopParams_.gauge` is the GLOBAL (replicated) host gauge field, `group` is the group of time ranks
        self.tsplit = tsplit
        ev0 = eigsolve_.eVecs[0]
        self.device = torch.device(device) if device is not None else (
            ev0.device if ev0.is_cuda else torch.device("cuda", torch.cuda.current_device()))
        self.dtype = eigsolve_.dtype
        if self.dtype not in (torch.complex128, torch.complex64):
            raise MugiqError("Loop_Mugiq: Precision not supported!")
        self.precision = PREC_DOUBLE if self.dtype == torch.complex128 else PREC_SINGLE
        self.L = eigsolve_.L
        self.lat = Lattice(self.L)
        if tsplit is not None:
            if tuple(tsplit.L_loc) != tuple(self.L):
                raise MugiqError(f"Loop_Mugiq: eigenvectors live on {self.L}, the T split expects {tsplit.L_loc}")
            if tsplit.world > 1 and group is None:
                raise MugiqError("Loop_Mugiq: a T split over several ranks needs the process group of the time ranks")
            self.cPrm = LoopComputeParam(loopParams_, self.L, comm_dim=(1, 1, 1, tsplit.world))
            tmax = max([b for (d, _, _, b) in self.cPrm.entries() if d == 3] + [0]) if self.cPrm.doNonLocal else 0
            if tmax > tsplit.H:
                raise MugiqError(f"Loop_Mugiq: t-displacements up to {tmax} need a halo of {tmax}, the T split has {tsplit.H}")
            self.L_run = tuple(tsplit.L_ext)  # the kernels run on the slab extended by the halos
        else:
            self.cPrm = LoopComputeParam(loopParams_, self.L)
            self.L_run = self.L
        self.writeDataPos = bool(loopParams_.writePosSpaceHDF5)
        self.writeDataMom = bool(loopParams_.writeMomSpaceHDF5)
        self.momSpaceFilename = loopParams_.fname_mom_h5
        self.posSpaceFilename = loopParams_.fname_pos_h5
        self.MomProjDone = False
        self.displace: Optional[Displace] = None
        self.dataPos = self.dataMom = self.dataMom_h = self.dataMom_bcast = None
        self.dataPosMP_d = self.dataMom_d = self.phaseMatrix_d = None
        self._allocateDataMemory()
        if self.cPrm.doMomProj:
            self._createPhaseMatrix()
        if self.cPrm.doNonLocal:
            if tsplit is not None:  # replicated global links -> this rank's extended slab
                import copy
                lp = copy.copy(loopParams_)
                if isinstance(loopParams_.gauge[0], torch.Tensor):  # this rank's extended slab, already on the device
                    if loopParams_.gauge[0].shape[0] != Lattice(self.L_run).volume:
                        raise MugiqError("Loop_Mugiq: device links given to a T split must be the rank's extended slab")
                else:
                    lp.gauge = [tsplit.global_slab(np.asarray(loopParams_.gauge[mu]), site_dim=0) for mu in range(4)]
                self.displace = Displace(lp, self.L_run, dtype=self.dtype, device=self.device)
            else:
                self.displace = Displace(loopParams_, self.L, dtype=self.dtype, device=self.device)
        self._plan = None
        self._plan_version = -1
        self._prepared = {}
        import os
        self._trace_on = os.environ.get("MUGIQ_B200_TSPLIT_TRACE", "") not in ("", "0")
        self._trace = []
        self._mp_workspace = None
        if self.cPrm.doMomProj and self.fused_momproj:
            need = max(ops.momproj_pos_workspace_bytes(LL, self.precision, self.cPrm.nLoop, self.cPrm.Nmom)
                       for LL in {tuple(self.L), tuple(self.L_run)})
            self._mp_workspace = torch.empty(need, dtype=torch.uint8, device=self.device)

    # -- lib/loop_mugiq.cpp:102-158 ---------------------------------------------------------------------
    def _allocateDataMemory(self):
        p = self.cPrm
        self.nElemPosLocPerLoop = p.nG * p.locV4
        self.nElemPosLoc = self.nElemPosLocPerLoop * p.nLoop
        self.nElemMomLoc = p.nG * p.Nmom * p.locT * p.nLoop
        self.nElemMomTot = p.nG * p.Nmom * p.totT * p.nLoop
        self.nElemPhMat = p.Nmom * p.locV3
        if self.peer_reduce:  # one IPC-shareable allocation the other ranks map
            self._pos_buf = ops.PeerBuffer(p.nLoop * p.nG * p.locV4 * (16 if self.dtype == torch.complex128 else 8), device=self.device)
            self.dataPos_d = self._pos_buf.tensor((p.nLoop, p.nG, p.locV4), self.dtype)
            self.dataPos_d.zero_()
        else:
            self.dataPos_d = torch.zeros((p.nLoop, p.nG, p.locV4), dtype=self.dtype, device=self.device)
        self.dataPosExt_d = None
        if self.tsplit is not None:
            self.dataPosExt_d = torch.zeros((p.nLoop, p.nG, Lattice(self.L_run).volume), dtype=self.dtype, device=self.device)
        if p.doMomProj and not self.fused_momproj:
            self.dataPosMP_d = torch.zeros((p.locV3, p.nData, p.locT), dtype=self.dtype, device=self.device)

    def _createPhaseMatrix(self):
        p = self.cPrm
        mom = np.asarray(p.momMatrix, dtype=np.int32).reshape(p.Nmom, 3)
        if self.fused_momproj:
            self.phaseMatrix_d = ops.phase_matrix_eo(mom, p.FTSign, p.localL, p.totalL, (0, 0, 0, 0), dtype=self.dtype,
                                                     device=self.device)
        else:
            self.phaseMatrix_d = ops.phase_matrix(mom, p.FTSign, p.localL, p.totalL, (0, 0, 0, 0), dtype=self.dtype,
                                                  device=self.device)

    # -- lib/loop_mugiq.cpp:440-525 ---------------------------------------------------------------------
    def computeCoarseLoop(self):
        p = self.cPrm
        es = self.eigsolve
        with torch.cuda.device(self.device):
            plan = self._loop_plan()
            if self.tsplit is not None:
                self._accumulate_tsplit(plan)
                return self._finish_tsplit()
            # eigenvector shards: the position-space buffer is summed over the group only when it is
            # needed on every rank (host copy requested or no momentum projection); otherwise the
            # projection, which is linear, runs on the partial sums and the small dataMom is reduced
            sharded = self.group is not None or self.comm is not None
            want_pos = self.reduce_pos if self.reduce_pos is not None else (self.copy_pos_to_host or not p.doMomProj)
            self._reduce_mom = sharded and p.doMomProj and not want_pos
            reduce_pos = sharded and not self._reduce_mom
            fused_reduce = reduce_pos and self.comm is not None   # sum overlapped with the kernels of the last batch
            if es.eVecs[0].is_cuda:
                starts = list(range(0, es.nEv, self.evec_batch))
                for b0 in starts:
                    b1 = min(es.nEv, b0 + self.evec_batch)
                    prep = self._prepared_batch(plan, b0, b1, es.eVecs[b0:b1])
                    if fused_reduce and b0 == starts[-1]:
                        if self.peer_reduce and self._peer is None:
                            self._attach_peer_reduce(plan)
                        plan.accumulate_allreduce(self.dataPos_d, prep, self.comm, accumulate=b0 > 0,
                                                  nchunks=self.allreduce_chunks)
                    else:
                        plan.accumulate(self.dataPos_d, prep, accumulate=b0 > 0)
            else:
                self._accumulate_from_host(plan)
                if fused_reduce:
                    self.comm.allreduce_pos(self.dataPos_d, plan.computed_slots(), self.L)
            if reduce_pos and not fused_reduce:
                allreduce_loop_buffer(self.dataPos_d, group=self.group)
            # slots derived after the eigenvector sum (minus-direction partners, repeated entries): linear in the
            # computed slots, so they are filled after the cross-rank sum and need none of their own
            plan.finalize(self.dataPos_d)
            # "Always copy the device position-space buffer to the host" (lib/loop_mugiq.cpp:512)
            if self.copy_pos_to_host:
                if self.dataPos is None:
                    self.dataPos = torch.empty(self.dataPos_d.shape, dtype=self.dtype, pin_memory=True)
                self.dataPos.copy_(self.dataPos_d, non_blocking=True)
                torch.cuda.current_stream().synchronize()
            if p.doMomProj:
                self.performMomentumProjection()
        return self

    def _attach_peer_reduce(self, plan):
        """Peer transport of the overlapped sum: exchange the IPC handles of dataPos_d and of a staging area over the process
        group, map every peer's pair, hand the tables to the communicator (once per Loop_Mugiq)."""
        import torch.distributed as dist
        comm = self.comm
        nbytes = comm.stage_bytes(plan, self.allreduce_chunks)
        stage = ops.PeerBuffer(nbytes, device=self.device)
        mine = (self._pos_buf.handle, stage.handle)
        every = [None] * comm.size
        dist.all_gather_object(every, mine, group=self.group)
        pos_ptrs, stage_ptrs, opened = [], [], []
        for r, (hp, hs) in enumerate(every):
            if r == comm.rank:
                pos_ptrs.append(self._pos_buf.ptr)
                stage_ptrs.append(stage.ptr)
            else:
                a, b = ops.peer_open(hp, self.device), ops.peer_open(hs, self.device)
                opened += [a, b]
                pos_ptrs.append(a)
                stage_ptrs.append(b)
        comm.attach_peers(pos_ptrs, stage_ptrs, nbytes)
        self._peer = {"stage": stage, "opened": opened}

    def close_peer_reduce(self):
        """Collective, last call on a Loop_Mugiq made with peer_reduce=True: detach and unmap the peer transport, then free
        the staging area and the shared position-space buffer (dataPos_d is gone afterwards; the host copies stay)."""
        if not self.peer_reduce:
            return
        import torch.distributed as dist
        torch.cuda.synchronize(self.device)
        if self._peer is not None:
            self.comm.attach_peers(None, None, 0)
            dist.barrier(group=self.group)
            for ptr in self._peer["opened"]:
                ops.peer_close(ptr, self.device)
            dist.barrier(group=self.group)
            self._peer["stage"].free()
            self._peer = None
        self.dataPos_d = None
        self._pos_buf.free()
        self.peer_reduce = False

    def _prepared_batch(self, plan, b0, b1, vecs):
        """Argument tables (pointer array, sigma array) of the resident batch [b0, b1): built once and reused while the
        batch consists of the same device buffers AND the same sigma values - the key holds every pointer and every
        sigma, so refilled buffers keep their table but a changed eVals_sigma or a swapped tensor rebuilds it (the
        reference re-reads sigma on every call, lib/loop_mugiq.cpp:479)."""
        # (this runs before the first kernel of a step can be launched: C-level map + one tobytes, ~20 us for 200 vectors)
        key = (tuple(map(torch.Tensor.data_ptr, vecs)), np.asarray(self.eigsolve.eVals_sigma[b0:b1], dtype=np.float64).tobytes())
        prep = self._prepared.get((b0, b1))
        if prep is None or prep[0] != key:
            prep = (key, plan.prepare(vecs, self.eigsolve.eVals_sigma[b0:b1]))
            self._prepared[(b0, b1)] = prep
        return prep[1]

    def _mark(self, name):
        """MUGIQ_B200_TSPLIT_TRACE=1: CUDA-event + host time stamps at the phase boundaries of a T-split step."""
        if not self._trace_on:
            return
        import time
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self._trace.append((name, ev, time.perf_counter()))

    def trace_report(self):
        """[(phase, device ms since the previous mark, host ms since the previous mark)] of the last traced step."""
        out = []
        for (n0, e0, h0), (n1, e1, h1) in zip(self._trace[:-1], self._trace[1:]):
            out.append((n1, e0.elapsed_time(e1), (h1 - h0) * 1e3))
        return out

    def _accumulate_tsplit(self, plan):
        """T split: every eigenvector batch is extended by the neighbours' halo slices (NCCL P2P over NVLink), then
        the fused kernel runs on the extended slab."""
        es, ts = self.eigsolve, self.tsplit
        nb = max(1, min(self.stream_batch, es.nEv))
        # the halo of the first batch is the only one nothing hides (0.29 ms for 200 slabs of 16^3x32): start with an eighth
        # of a batch, whose halo lands ~8x sooner, and push the rest under its kernels
        # (the first regular batch is split in two, so every later batch starts where it did before)
        batches = [(b0, min(es.nEv, b0 + nb)) for b0 in range(0, es.nEv, nb)]
        first = max(8, nb // 8)
        if batches[0][1] - batches[0][0] > 2 * first:
            batches = [(0, first), (first, batches[0][1])] + batches[1:]
        # only the interior is computed; the halos are read.  Plus-t loops read eigenvector slices above the interior,
        # minus-t loops that are computed directly read below it; minus-t loops DERIVED from their plus partner need the
        # partner's loop values below the interior instead, fetched once after the eigenvector sum.
        self._trace = []
        self._mark("start")
        plan.set_t_range(ts.H, ts.H + ts.Tl)
        lo, up, ll = plan.t_halo()
        self.tsplit_halo_sides = int(lo > 0) + int(up > 0)
        self.tsplit_halo_slices = lo + up  # only the slices that are read travel (H is rounded up to even)
        ext_kw = dict(group=self.group, device=self.device, lower=lo, upper=up)
        # the halo exchange of batch i+1 is posted before the kernels of batch i are launched, so it overlaps them
        pending = ts.begin_extend(es.eVecs[batches[0][0]:batches[0][1]], **ext_kw)
        for i, (b0, b1) in enumerate(batches):
            cur = pending
            if i + 1 < len(batches):
                n0, n1 = batches[i + 1]
                pending = ts.begin_extend(es.eVecs[n0:n1], **ext_kw)
            ext = ts.finish_extend(cur)
            self._mark(f"halo wait {i}")
            if ts.peer is not None:  # the slabs are extended in place: same device buffers every time
                plan.accumulate(self.dataPosExt_d, self._prepared_batch(plan, b0, b1, list(ext)), accumulate=b0 > 0)
            else:
                plan.accumulate(self.dataPosExt_d, list(ext), es.eVals_sigma[b0:b1], accumulate=b0 > 0)
            self._mark(f"kernels {i}")
        if ll > 0:
            slots, iL = [], 1
            for (d, sgn, a, b) in self.cPrm.entries():
                if d == 3 and sgn == 1:
                    slots += list(range(iL, iL + b - a + 1))
                iL += b - a + 1
            ts.exchange_loop_halo(self.dataPosExt_d, slots, group=self.group, depth=ll)
            self._mark("loop halo")
        plan.finalize(self.dataPosExt_d)
        self._mark("finalize")

    def _finish_tsplit(self):
        p, ts = self.cPrm, self.tsplit
        # every rank owns its time-slices: no reduction, only the interior of the extended buffer is kept
        # the interior of the extended buffer is the rank's position-space result; when only momentum-space data are
        # wanted (no host copy of dataPos) the projection reads the extended buffer in place and the halo time-slices of
        # its small result are dropped instead (dataPos_interior() still materialises the interior on demand)
        self._project_ext = p.doMomProj and self.fused_momproj and not self.copy_pos_to_host
        self._pos_valid = not self._project_ext
        if not self._project_ext:
            self.dataPos_d.copy_(ts.interior(self.dataPosExt_d, site_dim=2))
        self._mark("interior copy")
        self._reduce_mom = False
        if self.copy_pos_to_host:
            if self.dataPos is None:
                self.dataPos = torch.empty(self.dataPos_d.shape, dtype=self.dtype, pin_memory=True)
            self.dataPos.copy_(self.dataPos_d, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        if p.doMomProj:
            self.performMomentumProjection()
            self._mark("projection + gather + D2H")
        return self

    def dataPos_interior(self):
        """Device position-space buffer of this rank's own time-slices (T split: cut out of the extended buffer when the
        last run skipped that copy)."""
        if self.tsplit is not None and not getattr(self, "_pos_valid", True):
            self.dataPos_d.copy_(self.tsplit.interior(self.dataPosExt_d, site_dim=2))
            self._pos_valid = True
        return self.dataPos_d

    def _loop_plan(self):
        """The Wilson lines + launch schedule for (gauge field, entries): built once, rebuilt when the gauge field
        is uploaded again (Displace.upload_gauge bumps gaugeVersion)."""
        p = self.cPrm
        version = self.displace.gaugeVersion if self.displace is not None else 0
        if self._plan is None or self._plan_version != version:
            if self._plan is not None:
                self._plan.close()
            entries = p.entries() if p.doNonLocal else []
            gauge = self.displace.gaugeField if self.displace is not None else None
            self._plan = ops.LoopPlan(gauge, entries, self.L_run, self.precision)
            if self.eigsolve.field_order:
                if self.tsplit is not None:
                    raise MugiqError("Loop_Mugiq: the lattice-T split takes site-major eigenvectors")
                self._plan.set_evec_order(self.eigsolve.field_order)
            self._plan_version = version
        return self._plan

    def _accumulate_from_host(self, plan):
        """Eigenvectors resident in (pinned) HOST memory: the library's streamed feed (mugiq_b200_loop_feed_*) moves them
        batch by batch into a ring of device staging batches on its copy stream - consecutive host fields as ONE copy -
        while the loop kernels of the previous batch run; what the reference does per eigenvector with
        `*fineEvecL = *eVecs[n]` / prolongateEvec (lib/loop_mugiq.cpp:478-483)."""
        es = self.eigsolve
        nb = max(1, min(self.stream_batch, es.nEv))
        if getattr(self, "_feed", None) is None or self._feed.batch != nb:
            if getattr(self, "_feed", None) is not None:
                self._feed.close()
            self._feed = ops.LoopFeed(plan, self.dataPos_d, batch=nb, nbuf=2, order=es.field_order)
        elif self._feed.plan is not plan or self._feed.dataPos is not self.dataPos_d:
            self._feed.set_plan(plan, self.dataPos_d)   # the plan was rebuilt for a new gauge field
        self._feed.push_host(es.eVecs, es.eVals_sigma)
        self._feed.finish()

    # -- lib/loop_mugiq.cpp:323-434 ---------------------------------------------------------------------
    def performMomentumProjection(self):
        if self.MomProjDone:
            raise MugiqError("performMomentumProjection: Not supposed to be called more than once!!")
        p = self.cPrm
        if p.nData != p.nLoop * 16:
            raise MugiqError("performMomentumProjection: This function assumes that nData = nLoop * NGamma")
        if self.fused_momproj and getattr(self, "_project_ext", False):
            ts = self.tsplit  # H is even: a site has the same parity in the extended and in the local lattice
            ext = ops.momproj_pos(self.dataPosExt_d, self.phaseMatrix_d, p.nLoop, self.L_run, workspace=self._mp_workspace)
            self.dataMom_d = ext[:, :, ts.H:ts.H + ts.Tl].contiguous()
        elif self.fused_momproj:
            self.dataMom_d = ops.momproj_pos(self.dataPos_d, self.phaseMatrix_d, p.nLoop, self.L, workspace=self._mp_workspace)
        else:
            ops.reorder_mapgamma(self.dataPosMP_d, self.dataPos_d, p.nData, p.nLoop, self.L)
            M, N, K = p.locT * p.nData, p.Nmom, p.locV3
            self.dataMom_d = ops.momproj(self.dataPosMP_d, self.phaseMatrix_d, M, N, K).reshape(p.Nmom, p.nData, p.locT)
        if getattr(self, "_reduce_mom", False):
            if self.comm is not None:
                self.comm.allreduce(self.dataMom_d)
            else:
                allreduce_loop_buffer(self.dataMom_d, group=self.group)
        self.dataMom_h = self.dataMom_d.cpu()
        # single spatial block: MPI_Reduce over COMM_SPACE is the identity
        self.dataMom = self.dataMom_h
        if self.tsplit is not None:  # MPI_Gather over COMM_TIME + MPI_Bcast (lib/loop_mugiq.cpp:420-424)
            self.dataMom_bcast = self.tsplit.gather_time(self.dataMom_d, group=self.group).cpu()
        else:
            self.dataMom_bcast = self.dataMom_h
        self.MomProjDone = True

    # -- result access in the reference's HDF5 naming (lib/loop_mugiq.cpp:582-633) -------------------------
    def momentum_loops(self):
        """{(mom tuple, disp tag, gamma name): complex array [T]} — the datasets writeLoopsHDF5_Mom creates."""
        if not self.MomProjDone:
            raise MugiqError("momentum_loops: momentum projection has not been performed")
        p = self.cPrm
        mom = np.asarray(p.momMatrix).reshape(p.Nmom, 3)
        out = {}
        data = self.dataMom.numpy()
        for im in range(p.Nmom):
            for iL, tag in enumerate(p.loop_tags()):
                for ig in range(16):
                    out[(tuple(int(x) for x in mom[im]), tag, GAMMA_NAMES[ig])] = data[im, ig + 16 * iL, :]
        return out

    def writeLoopsHDF5(self):
        """Flag handling of lib/loop_mugiq.cpp:669-693; the file itself is written by mugiq_b200.h5lite."""
        p = self.cPrm
        if p.doMomProj:
            if not self.writeDataMom:
                warnings.warn("writeLoopsHDF5: Performed momentum projection, but got writeDatMom = FALSE. "
                              "Will proceed to write momentum-space loop data")
                self.writeDataMom = True
            from .h5lite import write_momentum_loops, write_momentum_loops_time_ranks
            ts = self.tsplit
            if ts is not None and ts.world > 1 and str(self.momSpaceFilename).lower().endswith((".h5", ".hdf5")):
                # every time rank writes its own rows of every [totT][2] dataset (lib/loop_mugiq.cpp:561-572,624)
                import torch.distributed as dist
                write_momentum_loops_time_ranks(self.momSpaceFilename, self, ts.rank, ts.world, lambda: dist.barrier(group=self.group))
            elif ts is None or ts.rank == 0:
                write_momentum_loops(self.momSpaceFilename, self)
        elif not self.writeDataPos:
            warnings.warn("writeLoopsHDF5: Did not perform momentum projection, but got writeDatPos = FALSE. "
                          "Will proceed to write position-space loop data")
            self.writeDataPos = True
        if self.writeDataPos:
            raise MugiqError("writeLoopsHDF5_Pos: Not supported yet!")  # lib/loop_mugiq.cpp:661-663


def computeLoop(loopParams: MugiqLoopParam, eigsolve: Eigsolve, **kw):
    """computeLoop<Float,order>(MugiqLoopParam, Eigsolve_Mugiq*) (lib/interface_mugiq.cpp:158-172)."""
    loop = Loop_Mugiq(loopParams, eigsolve, **kw)
    loop.computeCoarseLoop()
    if loopParams.writeMomSpaceHDF5 or loopParams.writePosSpaceHDF5:
        loop.writeLoopsHDF5()
    else:
        warnings.warn("computeLoop: Will NOT write output data!")
    return loop
