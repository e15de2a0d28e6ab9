"""ctypes front end of oracle/liboracle_mugiq.so (the CPU restatement in mugiq_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  Pinned by golden vectors of the reference's own CUDA kernels (tests/golden/ref_kernels_*.npz,
oracle/ref_kernels.py); QUDA's accessor conventions remain assumptions (see the header of mugiq_oracle.cpp).
All arrays are numpy, complex128 or complex64, in the same memory orders the C-ABI uses."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle_mugiq.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_cb_index.restype = C.c_int
    return _lib


def _suf(a):
    if a.dtype == np.complex128:
        return "f64"
    if a.dtype == np.complex64:
        return "f32"
    raise TypeError(a.dtype)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _i4(L):
    return (C.c_int * 4)(*[int(x) for x in L])


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def gamma_tables():
    rv = np.zeros((16, 4, 2)); ci = np.zeros((16, 4), dtype=np.int32)
    ms = np.zeros(16); mi = np.zeros(16, dtype=np.int32)
    lib().orc_gamma_tables(_p(rv), _p(ci), _p(ms), _p(mi))
    return rv, ci, ms, mi


def get_coords(cb, L, parity):
    x = (C.c_int * 4)()
    lib().orc_get_coords(x, int(cb), _i4(L), int(parity))
    return list(x)


def cb_index(x, L):
    return lib().orc_cb_index(_i4(x), _i4(L))


def contract(loop, vL, vR, sigma, L):
    """loop [16, V4] += (1/sigma) Tr[vL^dag Gamma vR]  (in place)."""
    getattr(lib(), "orc_contract_" + _suf(loop))(_p(loop), _p(vL), _p(vR), C.c_double(sigma), _i4(L))
    return loop


def displace(src, gauge, direction, sign, L):
    dst = np.empty_like(src)
    getattr(lib(), "orc_displace_" + _suf(src))(_p(dst), _p(src), _p(gauge), int(direction), int(sign), _i4(L))
    return dst


def reorder_mapgamma(inp, nLoop, L):
    """inp [nLoop,16,V4] -> out [V3, 16*nLoop, Lt]."""
    V3, Lt = int(L[0]) * int(L[1]) * int(L[2]), int(L[3])
    out = np.empty((V3, 16 * nLoop, Lt), dtype=inp.dtype)
    getattr(lib(), "orc_reorder_mapgamma_" + _suf(inp))(_p(out), _p(inp), int(nLoop), _i4(L))
    return out


def phase_matrix(mom, ftsign, localL, totalL=None, commCoord=(0, 0, 0, 0), dtype=np.complex128):
    mom = np.ascontiguousarray(np.asarray(mom, dtype=np.int32).reshape(-1, 3))
    totalL = totalL or localL
    V3 = int(localL[0]) * int(localL[1]) * int(localL[2])
    out = np.empty((mom.shape[0], V3), dtype=dtype)
    getattr(lib(), "orc_phase_matrix_" + _suf(out))(_p(out), _p(mom), mom.shape[0], int(ftsign), _i4(localL),
                                                    _i4(totalL), _i4(commCoord))
    return out


def gemm(A, B, M, N, K):
    """column-major C(MxN) = A(MxK) B(KxN); A given as any array whose memory is m + M*k, B as k + K*n."""
    out = np.empty((N, M), dtype=A.dtype)
    getattr(lib(), "orc_gemm_" + _suf(A))(_p(out), _p(A), _p(B), C.c_longlong(M), int(N), C.c_longlong(K))
    return out


def compute_loop(evecs, sigma, gauge, entries, L):
    """Loop_Mugiq::computeCoarseLoop in the reference's schedule.  evecs [nEv, V4, 12]; entries list of
    (dir, sign, start, stop).  Returns dataPos [nLoop, 16, V4]."""
    evecs = np.ascontiguousarray(evecs)
    nEv = evecs.shape[0]
    V4 = evecs.shape[1]
    nLoop = 1 + sum(b - a + 1 for (_, _, a, b) in entries)
    out = np.zeros((nLoop, 16, V4), dtype=evecs.dtype)
    ne = len(entries)
    arr = lambda k: (C.c_int * max(ne, 1))(*[int(e[k]) for e in entries])
    sig = (C.c_double * nEv)(*[float(s) for s in sigma])
    stride = V4 * 12 * 2
    getattr(lib(), "orc_compute_loop_" + _suf(evecs))(
        _p(out), _p(evecs), C.c_longlong(stride), sig, nEv, _p(gauge) if gauge is not None else None, ne,
        arr(0), arr(1), arr(2), arr(3), _i4(L))
    return out
