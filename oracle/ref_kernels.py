"""ctypes front end of oracle/_ref/libmugiq_ref.so: the reference's OWN CUDA wrappers and kernels
(/root/reference/lib/contract_wrappers.cu, lib/mugiq_{contract,displace,util}_kernels.cu), compiled unmodified against
oracle/quda_shim/ by `make -C oracle ref` (needs /root/reference; the built library travels to the GPU box).

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's reference legs).  Operands are torch CUDA tensors.
Colour-spinor fields are in QUDA's native orders (FLOAT2 = 2, FLOAT4 = 4, see quda_shim_core.h); site_to_quda /
quda_to_site convert from / to the canonical site-major order with plain numpy-style indexing (independent of the
product's conversion kernels).  The gauge field is [dir][parity][x_cb][3][3], the host QDP order."""
import ctypes as C
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libmugiq_ref.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def build(reference="/root/reference"):
    """Compiles the reference's kernel sources where they lie (only possible where /root/reference exists)."""
    if not os.path.isdir(reference):
        return False
    subprocess.check_call(["make", "-s", "-C", _HERE, "ref", f"REF={reference}"])
    return True


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libmugiq_ref.so is missing: run `make -C oracle ref` where /root/reference exists")
        _lib = C.CDLL(LIB_PATH)
    return _lib


def _i4(L):
    return (C.c_int * 4)(*[int(x) for x in L])


def _prec(t):
    return 8 if t.dtype == torch.complex128 else 4


def _p(t):
    return C.c_void_p(t.data_ptr())


def site_to_quda(v, order):
    """[V4, 12] site-major -> QUDA FLOAT2 ([parity][12][x_cb]) or FLOAT4 ([parity][6][x_cb][2])."""
    V4 = v.shape[0]
    a = v.reshape(2, V4 // 2, 12)
    if order == 2:
        return a.permute(0, 2, 1).contiguous().reshape(V4, 12)
    return a.reshape(2, V4 // 2, 6, 2).permute(0, 2, 1, 3).contiguous().reshape(V4, 12)


def quda_to_site(q, order):
    V4 = q.shape[0]
    if order == 2:
        return q.reshape(2, 12, V4 // 2).permute(0, 2, 1).contiguous().reshape(V4, 12)
    return q.reshape(2, 6, V4 // 2, 2).permute(0, 2, 1, 3).contiguous().reshape(V4, 12)


def _ok(rc, who):
    torch.cuda.synchronize()
    if rc != 0:
        raise RuntimeError(f"{who}: CUDA error in the reference path")


def contract(loop, vL_q, vR_q, sigma, L, order=2, quiet=True):
    """performLoopContraction: loop [16, V4] += (1/sigma) Tr[vL^dag Gamma vR] (in place, synchronous)."""
    torch.cuda.synchronize()
    _ok(lib().mugiq_ref_contract(_p(loop), _p(vL_q), _p(vR_q), C.c_double(float(sigma)), _i4(L), _prec(loop), int(order),
                                 int(bool(quiet))), "mugiq_ref_contract")
    return loop


def displace(dst_q, src_q, gauge, direction, sign, L, order=2, extended=True):
    """performCovariantDisplacementVector (dst != src)."""
    torch.cuda.synchronize()
    _ok(lib().mugiq_ref_displace(_p(dst_q), _p(src_q), _p(gauge), int(direction), int(sign), _i4(L), _prec(src_q), int(order),
                                 int(bool(extended))), "mugiq_ref_displace")
    return dst_q


def reorder_mapgamma(out, inp, nLoop, L):
    """convertIdxOrder_mapGamma: out [V3, 16*nLoop, Lt] from inp [nLoop, 16, V4]."""
    torch.cuda.synchronize()
    _ok(lib().mugiq_ref_reorder(_p(out), _p(inp), int(nLoop), _i4(L), _prec(inp)), "mugiq_ref_reorder")
    return out


def phase_matrix(mom, ftsign, L, totalL=None, dtype=torch.complex128):
    """createPhaseMatrixGPU: returns phase [Nmom, V3]."""
    import numpy as np
    mom = np.ascontiguousarray(np.asarray(mom, dtype=np.int32).reshape(-1, 3))
    out = torch.empty((mom.shape[0], int(L[0]) * int(L[1]) * int(L[2])), dtype=dtype, device="cuda")
    torch.cuda.synchronize()
    _ok(lib().mugiq_ref_phase(_p(out), mom.ctypes.data_as(C.c_void_p), mom.shape[0], int(ftsign), _i4(L), _i4(totalL or L),
                              _prec(out)), "mugiq_ref_phase")
    return out


def compute_loop(evecs_q, sigma, gauge, entries, L, order=2):
    """The eigenvector x displacement loop nest of Loop_Mugiq::computeCoarseLoop around the reference's wrappers
    (FP64).  entries: [(dir, sign, start, stop)].  Returns dataPos [nLoop, 16, V4]."""
    V4 = int(L[0]) * int(L[1]) * int(L[2]) * int(L[3])
    nLoop = 1 + sum(b - a + 1 for (_, _, a, b) in entries)
    pos = torch.zeros((nLoop, 16, V4), dtype=torch.complex128, device="cuda")
    work = torch.empty((3, V4, 12), dtype=torch.complex128, device="cuda")
    n = len(evecs_q)
    ptrs = (C.c_void_p * n)(*[v.data_ptr() for v in evecs_q])
    sig = (C.c_double * n)(*[float(s) for s in sigma])
    ent = (C.c_int * max(4 * len(entries), 1))(*[int(x) for e in entries for x in e])
    torch.cuda.synchronize()
    _ok(lib().mugiq_ref_loop(_p(pos), ptrs, sig, n, _p(gauge) if gauge is not None else None, ent, len(entries), _i4(L),
                             int(order), _p(work)), "mugiq_ref_loop")
    return pos


def kernel_only_ms(evecs_q, sigma, gauge, entries, L):
    """The reference's loop nest with CUDA events around its KERNELS only (argument structs pre-staged, no allocation or
    synchronisation in the timed regions): {"ms": contraction + displacement kernels, "contract_ms", "displace_ms",
    "field_copies_ms", "launches", "note"}.  FLOAT2 order, FP64."""
    V4 = int(L[0]) * int(L[1]) * int(L[2]) * int(L[3])
    nLoop = 1 + sum(b - a + 1 for (_, _, a, b) in entries)
    pos = torch.zeros((nLoop, 16, V4), dtype=torch.complex128, device="cuda")
    work = torch.empty((3, V4, 12), dtype=torch.complex128, device="cuda")
    n = len(evecs_q)
    ptrs = (C.c_void_p * n)(*[v.data_ptr() for v in evecs_q])
    sig = (C.c_double * n)(*[float(s) for s in sigma])
    ent = (C.c_int * max(4 * len(entries), 1))(*[int(x) for e in entries for x in e])
    ms = (C.c_float * 3)()
    cnt = (C.c_int * 3)()
    torch.cuda.synchronize()
    _ok(lib().mugiq_ref_loop_kernel_times(_p(pos), ptrs, sig, n, _p(gauge) if gauge is not None else None, ent, len(entries), _i4(L),
                                          _p(work), ms, cnt), "mugiq_ref_loop_kernel_times")
    return {"ms": ms[0] + ms[1], "contract_ms": ms[0], "displace_ms": ms[1], "field_copies_ms": ms[2],
            "launches": cnt[0] + cnt[1],
            "note": "CUDA events around loopContract_kernel / covariantDisplacementVector_kernel only (launch geometry of "
                    "lib/contract_wrappers.cu:103-110,186-191, argument structs staged outside the timed regions); the field "
                    "copies and zeroing of the loop nest are timed separately"}
