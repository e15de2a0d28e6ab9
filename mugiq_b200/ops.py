"""Thin torch-tensor front end of the C-ABI (include/mugiq_b200.h): one function per entry point, same
names and argument meaning.  Tensors only carry device memory and the current CUDA stream; every function
checks that its operands live on a CUDA device and raises otherwise (no CPU path)."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, make_geom, ptr_array, entry_array, PREC_DOUBLE, PREC_SINGLE


def _prec(t):
    if t.dtype in (torch.complex128, torch.float64):
        return PREC_DOUBLE
    if t.dtype in (torch.complex64, torch.float32):
        return PREC_SINGLE
    raise TypeError(f"unsupported dtype {t.dtype}")


def _dev(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("mugiq_b200 operates on CUDA tensors only (no CPU fallback)")
        if not t.is_contiguous():
            raise RuntimeError("mugiq_b200 needs contiguous tensors")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _volume(L):
    return int(L[0]) * int(L[1]) * int(L[2]) * int(L[3])


def gamma_tables():
    rv = np.zeros((16, 4, 2), dtype=np.float64)
    ci = np.zeros((16, 4), dtype=np.int32)
    ms = np.zeros(16, dtype=np.float64)
    mi = np.zeros(16, dtype=np.int32)
    check(_lib.load().mugiq_b200_gamma_tables(rv.ctypes.data_as(_lib._pd), ci.ctypes.data_as(_lib._pi),
                                              ms.ctypes.data_as(_lib._pd), mi.ctypes.data_as(_lib._pi)))
    return rv, ci, ms, mi


def device_info():
    name = C.create_string_buffer(128)
    sm, cc = C.c_int(), C.c_int()
    fr, tot = C.c_longlong(), C.c_longlong()
    check(_lib.load().mugiq_b200_device_info(name, 128, C.byref(sm), C.byref(cc), C.byref(fr), C.byref(tot)))
    return {"name": name.value.decode(), "sm_count": sm.value, "cc": cc.value, "free": fr.value, "total": tot.value}


def gauge_upload(gauge_h, L, device="cuda", out=None):
    """gauge_h: four host arrays (QDP order, one per direction) or one [4, volume, 3, 3] array."""
    dirs = [np.ascontiguousarray(gauge_h[mu]) for mu in range(4)]
    cdt = torch.complex128 if dirs[0].dtype == np.complex128 else torch.complex64
    if out is None:
        out = torch.empty((4, _volume(L), 3, 3), dtype=cdt, device=device)
    geom = make_geom(L, _prec(out))
    with torch.cuda.device(out.device):
        check(_lib.load().mugiq_b200_gauge_upload(out.data_ptr(), ptr_array([d.ctypes.data for d in dirs]),
                                                  C.byref(geom), _stream()))
    return out


def ingest_spinor(src, order, L):
    _dev(src)
    dst = torch.empty((_volume(L), 12), dtype=src.dtype, device=src.device)
    geom = make_geom(L, _prec(src))
    with torch.cuda.device(src.device):
        check(_lib.load().mugiq_b200_ingest_spinor(dst.data_ptr(), src.data_ptr(), order, C.byref(geom), _stream()))
    return dst


def ingest_spinor_batch(srcs, order, L, out=None):
    """QUDA-native fields -> one [n, V4, 12] site-major tensor (`out` if given), one launch."""
    _dev(*srcs)
    n = len(srcs)
    dst = out if out is not None else torch.empty((n, _volume(L), 12), dtype=srcs[0].dtype, device=srcs[0].device)
    _dev(dst)
    geom = make_geom(L, _prec(srcs[0]))
    with torch.cuda.device(dst.device):
        check(_lib.load().mugiq_b200_ingest_spinor_batch(ptr_array([dst[i].data_ptr() for i in range(n)]),
                                                         ptr_array([s.data_ptr() for s in srcs]), n, order, C.byref(geom),
                                                         _stream()))
    return dst


def export_spinor(src_site, order, L):
    _dev(src_site)
    dst = torch.empty_like(src_site)
    geom = make_geom(L, _prec(src_site))
    with torch.cuda.device(src_site.device):
        check(_lib.load().mugiq_b200_export_spinor(dst.data_ptr(), order, src_site.data_ptr(), C.byref(geom), _stream()))
    return dst


def contract(loop, vL, vR, sigma, L):
    """performLoopContraction: loop += (1/sigma) * Tr[vL^dag Gamma vR]."""
    _dev(loop, vL, vR)
    geom = make_geom(L, _prec(vL))
    with torch.cuda.device(vL.device):
        check(_lib.load().mugiq_b200_contract(loop.data_ptr(), vL.data_ptr(), vR.data_ptr(), float(sigma),
                                              C.byref(geom), _stream()))
    return loop


def contract_batch(loop, vLs, vRs, sigma, L, accumulate=True):
    _dev(loop, *vLs)
    if vRs is not None:
        _dev(*vRs)
    n = len(vLs)
    sig = (C.c_double * max(n, 1))(*[float(s) for s in sigma])
    geom = make_geom(L, _prec(loop))
    with torch.cuda.device(loop.device):
        check(_lib.load().mugiq_b200_contract_batch(
            loop.data_ptr(), ptr_array([v.data_ptr() for v in vLs]),
            ptr_array([v.data_ptr() for v in vRs]) if vRs is not None else None, sig, n, int(bool(accumulate)),
            C.byref(geom), _stream()))
    return loop


def displace(dst, src, gauge, direction, sign, L):
    """performCovariantDisplacementVector."""
    _dev(dst, src, gauge)
    geom = make_geom(L, _prec(src))
    with torch.cuda.device(src.device):
        check(_lib.load().mugiq_b200_displace(dst.data_ptr(), src.data_ptr(), gauge.data_ptr(), int(direction),
                                              int(sign), C.byref(geom), _stream()))
    return dst


def displace_batch(dsts, srcs, gauge, direction, sign, L):
    """One hop for a batch of fields in one launch (the link tile is loaded once per batch)."""
    _dev(gauge, *dsts, *srcs)
    n = len(srcs)
    geom = make_geom(L, _prec(srcs[0]))
    with torch.cuda.device(gauge.device):
        check(_lib.load().mugiq_b200_displace_batch(ptr_array([d.data_ptr() for d in dsts]),
                                                    ptr_array([s.data_ptr() for s in srcs]), n, gauge.data_ptr(),
                                                    int(direction), int(sign), C.byref(geom), _stream()))
    return dsts


def contract_native(loop, vLs, vRs, sigma, order, L, accumulate=True):
    """Contraction of fields in QUDA FLOAT2 (order 2) / FLOAT4 (order 4) order, no layout conversion."""
    _dev(loop, *vLs)
    if vRs is not None:
        _dev(*vRs)
    n = len(vLs)
    sig = (C.c_double * n)(*[float(s) for s in sigma])
    geom = make_geom(L, _prec(loop))
    with torch.cuda.device(loop.device):
        check(_lib.load().mugiq_b200_contract_native(
            loop.data_ptr(), ptr_array([v.data_ptr() for v in vLs]),
            ptr_array([v.data_ptr() for v in vRs]) if vRs is not None else None, sig, n, int(order), int(bool(accumulate)),
            C.byref(geom), _stream()))
    return loop


def displace_native(dsts, srcs, gauge, direction, sign, order, L):
    """One hop on fields in QUDA FLOAT2 / FLOAT4 order (source and destination)."""
    _dev(gauge, *dsts, *srcs)
    geom = make_geom(L, _prec(srcs[0]))
    with torch.cuda.device(gauge.device):
        check(_lib.load().mugiq_b200_displace_native(ptr_array([d.data_ptr() for d in dsts]),
                                                     ptr_array([s.data_ptr() for s in srcs]), len(srcs), gauge.data_ptr(),
                                                     int(direction), int(sign), int(order), C.byref(geom), _stream()))
    return dsts


def loop_workspace_bytes(L, precision, nvec, entries):
    geom = make_geom(L, precision)
    return check(_lib.load().mugiq_b200_loop_workspace_bytes(C.byref(geom), nvec, entry_array(entries), len(entries)))


def loop_accumulate(dataPos, evecs, sigma, gauge, entries, L, accumulate=False, workspace=None):
    """The eigenvector/displacement loop nest of Loop_Mugiq::computeCoarseLoop for the given eigenvectors."""
    _dev(dataPos, gauge, workspace, *evecs)
    n = len(evecs)
    prec = _prec(dataPos)
    geom = make_geom(L, prec)
    need = loop_workspace_bytes(L, prec, n, entries)
    if need > 0 and (workspace is None or workspace.numel() * workspace.element_size() < need):
        workspace = torch.empty(need, dtype=torch.uint8, device=dataPos.device)
    sig = (C.c_double * n)(*[float(s) for s in sigma])
    with torch.cuda.device(dataPos.device):
        check(_lib.load().mugiq_b200_loop_accumulate(
            dataPos.data_ptr(), ptr_array([v.data_ptr() for v in evecs]), sig, n,
            gauge.data_ptr() if gauge is not None else None, entry_array(entries), len(entries),
            int(bool(accumulate)), workspace.data_ptr() if workspace is not None else None, C.byref(geom), _stream()))
    return dataPos


class PreparedBatch:
    """Host-side argument tables of an eigenvector batch (LoopPlan.prepare)."""
    __slots__ = ("ptrs", "sigma", "n", "keep")

    def __init__(self, ptrs, sigma, n, keep):
        self.ptrs, self.sigma, self.n, self.keep = ptrs, sigma, n, keep


class LoopPlan:
    """mugiq_b200_loop_plan_*: owns the Wilson lines and the launch schedule for one (gauge field, entry list)."""

    def __init__(self, gauge, entries, L, precision=PREC_DOUBLE):
        self.L = tuple(int(x) for x in L)
        self.precision = precision
        self._gauge = gauge  # keep the device gauge field alive: one-link loops read it directly
        self._h = C.c_void_p()
        if gauge is not None:
            _dev(gauge)
        geom = make_geom(L, precision)
        dev = gauge.device if gauge is not None else torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(dev):
            check(_lib.load().mugiq_b200_loop_plan_create(C.byref(self._h), gauge.data_ptr() if gauge is not None else None,
                                                          entry_array(entries), len(entries), C.byref(geom), _stream()))
        self.nLoop = check(_lib.load().mugiq_b200_loop_plan_nloop(self._h))

    def info(self):
        a, b, c, w = C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
        check(_lib.load().mugiq_b200_loop_plan_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(w)))
        return {"computed": a.value, "derived": b.value, "groups": c.value, "wilson_bytes": w.value}

    def computed_slots(self):
        """dataPos slots accumulate() writes (the others are derived by finalize())."""
        n = check(_lib.load().mugiq_b200_loop_plan_computed_slots(self._h, None, 0))
        arr = (C.c_int * n)()
        check(_lib.load().mugiq_b200_loop_plan_computed_slots(self._h, arr, n))
        return list(arr)

    def set_evec_order(self, order):
        """0: canonical site-major eigenvectors (default); 2: QUDA FLOAT2 fields, staged by the fused kernel itself."""
        check(_lib.load().mugiq_b200_loop_plan_set_evec_order(self._h, int(order)))

    def set_t_range(self, t_begin, t_end):
        """accumulate() computes only the time-slices [t_begin, t_end) (interior of a lattice-T split slab)."""
        check(_lib.load().mugiq_b200_loop_plan_set_t_range(self._h, int(t_begin), int(t_end)))

    def t_halo(self):
        """(eigenvector slices read below the interior, above it, loop-buffer slices read below it)."""
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        check(_lib.load().mugiq_b200_loop_plan_t_halo(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def prepare(self, evecs, sigma):
        """Host-side argument tables of a batch (pointer array, sigma array): building them costs ~0.1 ms for 200
        eigenvectors, which a caller that runs the same resident batch again can pay once."""
        _dev(*evecs)
        n = len(evecs)
        return PreparedBatch(ptr_array([v.data_ptr() for v in evecs]), (C.c_double * n)(*[float(s) for s in sigma]), n,
                             list(evecs))  # the tensors are kept alive with the table

    def accumulate(self, dataPos, evecs, sigma=None, accumulate=False):
        """evecs: a sequence of eigenvector tensors (with `sigma`), or what prepare() returned for them."""
        prep = evecs if isinstance(evecs, PreparedBatch) else self.prepare(evecs, sigma)
        _dev(dataPos)
        with torch.cuda.device(dataPos.device):
            check(_lib.load().mugiq_b200_loop_plan_accumulate(self._h, dataPos.data_ptr(), prep.ptrs, prep.sigma, prep.n,
                                                              int(bool(accumulate)), _stream()))
        return dataPos

    def accumulate_allreduce(self, dataPos, evecs, comm, sigma=None, accumulate=False, nchunks=8):
        """accumulate() for the last (or only) batch of a rank's eigenvector shard, fused with the NCCL sum of every slot
        the plan computes: time-slice chunk k is summed on the communicator's side stream while chunk k+1 computes."""
        prep = evecs if isinstance(evecs, PreparedBatch) else self.prepare(evecs, sigma)
        _dev(dataPos)
        with torch.cuda.device(dataPos.device):
            check(_lib.load().mugiq_b200_loop_plan_accumulate_allreduce(self._h, dataPos.data_ptr(), prep.ptrs, prep.sigma, prep.n,
                                                                        int(bool(accumulate)), comm._h, int(nchunks), _stream()))
        return dataPos

    def finalize(self, dataPos, accumulate=False):
        _dev(dataPos)
        with torch.cuda.device(dataPos.device):
            check(_lib.load().mugiq_b200_loop_plan_finalize(self._h, dataPos.data_ptr(), int(bool(accumulate)), _stream()))
        return dataPos

    def close(self):
        if self._h:
            _lib.load().mugiq_b200_loop_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LoopFeed:
    """mugiq_b200_loop_feed_*: streamed eigenvector feed of a LoopPlan (device staging ring, producer / consumer streams)."""

    def __init__(self, plan, dataPos, batch=16, nbuf=2, order=0, accumulate=False):
        _dev(dataPos)
        self.plan, self.dataPos, self.batch = plan, dataPos, int(batch)
        self._h = C.c_void_p()
        with torch.cuda.device(dataPos.device):
            check(_lib.load().mugiq_b200_loop_feed_create(C.byref(self._h), plan._h, dataPos.data_ptr(), int(batch), int(nbuf),
                                                          int(order), int(bool(accumulate)), _stream()))

    def set_plan(self, plan, dataPos):
        """Point the feed at another plan / loop buffer of the same lattice (between runs); the staging ring is kept."""
        _dev(dataPos)
        check(_lib.load().mugiq_b200_loop_feed_set_plan(self._h, plan._h, dataPos.data_ptr()))
        self.plan, self.dataPos = plan, dataPos

    def push_host(self, evecs_h, sigma):
        """Host (pinned) eigenvector tensors -> staging batches (H2D on the feed's copy stream) -> loop kernels."""
        n = len(evecs_h)
        for v in evecs_h:
            if v.is_cuda or not v.is_contiguous():
                raise RuntimeError("LoopFeed.push_host takes contiguous host tensors")
        sig = (C.c_double * n)(*[float(s) for s in sigma])
        with torch.cuda.device(self.dataPos.device):
            check(_lib.load().mugiq_b200_loop_feed_push_host(self._h, ptr_array([v.data_ptr() for v in evecs_h]), sig, n))

    def acquire(self, n, producer_stream=None):
        """n device field pointers of the next staging batch, for a device-side producer working on `producer_stream`."""
        arr = (C.c_void_p * n)()
        st = C.c_void_p((producer_stream or torch.cuda.current_stream()).cuda_stream)
        with torch.cuda.device(self.dataPos.device):
            check(_lib.load().mugiq_b200_loop_feed_acquire(self._h, arr, int(n), st))
        return [int(p) for p in arr]

    def commit(self, sigma, producer_stream=None):
        n = len(sigma)
        sig = (C.c_double * n)(*[float(s) for s in sigma])
        st = C.c_void_p((producer_stream or torch.cuda.current_stream()).cuda_stream)
        with torch.cuda.device(self.dataPos.device):
            check(_lib.load().mugiq_b200_loop_feed_commit(self._h, sig, n, st))

    def finish(self):
        n = C.c_longlong()
        check(_lib.load().mugiq_b200_loop_feed_finish(self._h, C.byref(n)))
        return n.value

    def close(self):
        if self._h:
            with torch.cuda.device(self.dataPos.device):
                _lib.load().mugiq_b200_loop_feed_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def reorder_mapgamma(out, inp, nData, nLoop, L):
    _dev(out, inp)
    geom = make_geom(L, _prec(inp))
    with torch.cuda.device(inp.device):
        check(_lib.load().mugiq_b200_reorder_mapgamma(out.data_ptr(), inp.data_ptr(), int(nData), int(nLoop),
                                                      C.byref(geom), _stream()))
    return out


def phase_matrix(mom, ftsign, localL, totalL=None, commCoord=(0, 0, 0, 0), dtype=torch.complex128, device="cuda"):
    """mom: [Nmom][3] ints.  Returns phase[im, v3] (memory order v3 + V3*im)."""
    mom = np.ascontiguousarray(np.asarray(mom, dtype=np.int32).reshape(-1, 3))
    nmom = mom.shape[0]
    totalL = totalL or localL
    V3 = int(localL[0]) * int(localL[1]) * int(localL[2])
    out = torch.empty((nmom, V3), dtype=dtype, device=device)
    i4 = C.c_int * 4
    with torch.cuda.device(out.device):
        check(_lib.load().mugiq_b200_phase_matrix(out.data_ptr(), mom.ctypes.data_as(_lib._pi), nmom, int(ftsign),
                                                  i4(*localL), i4(*totalL), i4(*commCoord), _prec(out), _stream()))
    return out


def momproj(posMP, phase, M, N, K, workspace=None):
    """dataMom(M x N) = dataPosMP(M x K) * phase(K x N), column-major.  Returns a [N, M] tensor (memory m + M*n)."""
    _dev(posMP, phase, workspace)
    prec = _prec(posMP)
    need = check(_lib.load().mugiq_b200_momproj_workspace_bytes(M, N, K, prec))
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=posMP.device)
    out = torch.empty((N, M), dtype=posMP.dtype, device=posMP.device)
    with torch.cuda.device(posMP.device):
        check(_lib.load().mugiq_b200_momproj(out.data_ptr(), posMP.data_ptr(), phase.data_ptr(), M, N, K, prec,
                                             workspace.data_ptr(), _stream()))
    return out


def phase_matrix_eo(mom, ftsign, localL, totalL=None, commCoord=(0, 0, 0, 0), dtype=torch.complex128, device="cuda"):
    """Phase matrix in the even/odd run order momproj_pos reads: [2 (s = (t+parity)&1), Nmom, V3/2]."""
    mom = np.ascontiguousarray(np.asarray(mom, dtype=np.int32).reshape(-1, 3))
    nmom = mom.shape[0]
    totalL = totalL or localL
    V3h = int(localL[0]) * int(localL[1]) * int(localL[2]) // 2
    out = torch.empty((2, nmom, V3h), dtype=dtype, device=device)
    i4 = C.c_int * 4
    with torch.cuda.device(out.device):
        check(_lib.load().mugiq_b200_phase_matrix_eo(out.data_ptr(), mom.ctypes.data_as(_lib._pi), nmom, int(ftsign),
                                                     i4(*localL), i4(*totalL), i4(*commCoord), _prec(out), _stream()))
    return out


def momproj_pos_workspace_bytes(L, precision, nLoop, nmom):
    geom = make_geom(L, precision)
    return check(_lib.load().mugiq_b200_momproj_pos_workspace_bytes(C.byref(geom), int(nLoop), int(nmom)))


def momproj_pos(dataPos, phase_eo, nLoop, L, workspace=None):
    """Momentum projection straight from dataPos [nLoop,16,V4] (stages 3+4 fused).  Returns [Nmom, 16*nLoop, Lt]."""
    _dev(dataPos, phase_eo, workspace)
    prec = _prec(dataPos)
    nmom = phase_eo.shape[1]
    geom = make_geom(L, prec)
    need = check(_lib.load().mugiq_b200_momproj_pos_workspace_bytes(C.byref(geom), int(nLoop), int(nmom)))
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dataPos.device)
    out = torch.empty((nmom, 16 * int(nLoop), int(L[3])), dtype=dataPos.dtype, device=dataPos.device)
    with torch.cuda.device(dataPos.device):
        check(_lib.load().mugiq_b200_momproj_pos(out.data_ptr(), dataPos.data_ptr(), phase_eo.data_ptr(), int(nLoop), int(nmom),
                                                 C.byref(geom), workspace.data_ptr(), _stream()))
    return out


# ---- eigenvector shards: NCCL communicator of the library (mugiq_b200_comm_*, mugiq_b200_allreduce*) ----------------
class Comm:
    """mugiq_b200_comm_t over the ranks of a torch.distributed process group: rank 0 makes the NCCL id, the group
    carries its 128 bytes to the others, every rank joins on its current CUDA device.  The library binds the NCCL
    already loaded in the process (torch's), so there is one NCCL per process."""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        ident = C.create_string_buffer(_lib.COMM_ID_BYTES)
        if self.rank == 0:
            check(_lib.load().mugiq_b200_comm_unique_id(ident))
        box = [ident.raw]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(_lib.load().mugiq_b200_comm_create(C.byref(self._h), C.c_char_p(box[0]), self.rank, self.size))

    def info(self):
        r, n, v = C.c_int(), C.c_int(), C.c_int()
        check(_lib.load().mugiq_b200_comm_info(self._h, C.byref(r), C.byref(n), C.byref(v)))
        return {"rank": r.value, "size": n.value, "nccl_version": v.value}

    def allreduce(self, t):
        """In-place sum of a real or complex tensor over the ranks, on the current stream."""
        _dev(t)
        count = t.numel() * (2 if t.is_complex() else 1)
        with torch.cuda.device(t.device):
            check(_lib.load().mugiq_b200_allreduce(t.data_ptr(), count, _prec(t), self._h, _stream()))
        return t

    def allreduce_pos(self, dataPos, slots, L, t_begin=0, t_end=-1):
        """In-place sum of the time-slices [t_begin, t_end) of the loop slots `slots` of dataPos [nLoop, 16, V4]."""
        _dev(dataPos)
        geom = make_geom(L, _prec(dataPos))
        arr = (C.c_int * len(slots))(*[int(x) for x in slots])
        with torch.cuda.device(dataPos.device):
            check(_lib.load().mugiq_b200_allreduce_pos(dataPos.data_ptr(), arr, len(slots), int(t_begin), int(t_end),
                                                       C.byref(geom), self._h, _stream()))
        return dataPos

    def stage_bytes(self, plan, nchunks):
        """Staging bytes per rank of the peer transport for `plan` summed in `nchunks` chunks."""
        n = _lib.load().mugiq_b200_comm_stage_bytes(plan._h, int(nchunks), self.size)
        if n < 0:
            check(int(n))
        return int(n)

    def attach_peers(self, peer_pos, peer_stage, stage_bytes):
        """peer_pos / peer_stage: device addresses (ints) of every rank's position-space buffer / staging area as mapped in
        this process, own allocation at index `rank`; None, None detaches."""
        lib = _lib.load()
        if peer_pos is None:
            check(lib.mugiq_b200_comm_attach_peers(self._h, None, None, 0))
            return
        with torch.cuda.device(self.device):
            check(lib.mugiq_b200_comm_attach_peers(self._h, ptr_array(list(peer_pos)), ptr_array(list(peer_stage)), int(stage_bytes)))

    def allgather(self, send):
        _dev(send)
        recv = torch.empty((self.size,) + tuple(send.shape), dtype=send.dtype, device=send.device)
        with torch.cuda.device(send.device):
            check(_lib.load().mugiq_b200_allgather(recv.data_ptr(), send.data_ptr(), send.numel() * send.element_size(), self._h,
                                                   _stream()))
        return recv

    def close(self):
        if self._h:
            with torch.cuda.device(self.device):
                _lib.load().mugiq_b200_comm_destroy(self._h)
            self._h = C.c_void_p()


# ---- NVLink peer memory for the lattice-T split (mugiq_b200_peer_*, mugiq_b200_halo_push_t) -------------------------
class PeerBuffer:
    """One device allocation other processes can map (CUDA IPC): `handle` (64 bytes) goes to the peers, who call
    peer_open(handle).  tensor() views the allocation as a torch tensor without copying."""

    def __init__(self, nbytes, device=None):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        h = C.create_string_buffer(64)
        with torch.cuda.device(self.device):
            check(_lib.load().mugiq_b200_peer_alloc(C.byref(p), self.nbytes, h))
        self.ptr = p.value
        self.handle = h.raw

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 2}

    def tensor(self, shape, dtype):
        return torch.as_tensor(self, device=self.device).view(dtype).reshape(shape)

    def free(self):
        if self.ptr:
            with torch.cuda.device(self.device):
                check(_lib.load().mugiq_b200_peer_free(C.c_void_p(self.ptr)))
            self.ptr = None


def peer_open(handle, device=None):
    p = C.c_void_p()
    with torch.cuda.device(device if device is not None else torch.cuda.current_device()):
        check(_lib.load().mugiq_b200_peer_open(C.byref(p), C.c_char_p(handle)))
    return p.value


def peer_close(ptr, device=None):
    with torch.cuda.device(device if device is not None else torch.cuda.current_device()):
        check(_lib.load().mugiq_b200_peer_close(C.c_void_p(ptr)))


def halo_push_t(dst_ptr, src_ptr, first_vec, nvec, Lt_ext, V3h, src_t, dst_t, nslices, mode=0, site_bytes=192):
    """Time-slices [src_t, src_t+nslices) of every (vector, parity) block of a batch of extended slabs -> slices
    [dst_t, ...) of the same vectors in another (peer-mapped or local) allocation, on the current stream."""
    check(_lib.load().mugiq_b200_halo_push_t(C.c_void_p(dst_ptr), C.c_void_p(src_ptr), int(first_vec), int(nvec), int(Lt_ext),
                                             int(V3h), int(site_bytes), int(src_t), int(dst_t), int(nslices), int(mode), _stream()))


# ---- instrumentation (mugiq_b200_prof_*) ---------------------------------------------------------------
def prof_enable(on=True):
    check(_lib.load().mugiq_b200_prof_enable(int(bool(on))))


def prof_reset():
    check(_lib.load().mugiq_b200_prof_reset())


def prof_fused_trace(buf=None):
    """Per-CTA timeline of the following fused launches into `buf` (int64 CUDA tensor, 16 per CTA); None switches it off."""
    if buf is None:
        check(_lib.load().mugiq_b200_prof_fused_trace(None, 0))
    else:
        check(_lib.load().mugiq_b200_prof_fused_trace(C.c_void_p(buf.data_ptr()), buf.numel() // 16))


def prof_report():
    """{kernel name: {"launches", "timed", "ms", "alg_bytes"}} for every kernel launched since the last reset."""
    lib = _lib.load()
    out = {}
    for k in range(lib.mugiq_b200_prof_num_kernels()):
        n, t = C.c_longlong(), C.c_longlong()
        ms, b, f = C.c_double(), C.c_double(), C.c_double()
        check(lib.mugiq_b200_prof_query(k, C.byref(n), C.byref(t), C.byref(ms), C.byref(b), C.byref(f)))
        if n.value:
            out[lib.mugiq_b200_prof_name(k).decode()] = {"launches": n.value, "timed": t.value, "ms": ms.value,
                                                         "alg_bytes": b.value, "alg_flops": f.value}
    return out
