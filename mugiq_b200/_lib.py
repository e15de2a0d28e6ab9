"""ctypes binding of the C-ABI declared in include/mugiq_b200.h.

The shared library holds the hand-written sm_100a kernels; there is no CPU or PyTorch fallback: if the
library is missing, importing a compute entry point raises immediately.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmugiq_b200.so")

PREC_SINGLE, PREC_DOUBLE = 4, 8
ORDER_SITE, ORDER_FLOAT2, ORDER_FLOAT4 = 0, 2, 4
DIR_X, DIR_Y, DIR_Z, DIR_T = 0, 1, 2, 3
SIGN_MINUS, SIGN_PLUS = 0, 1
MAX_ENTRIES = 64
COMM_ID_BYTES = 128


class Geom(C.Structure):
    """mugiq_b200_geom_t"""
    _fields_ = [("L", C.c_int * 4), ("precision", C.c_int)]


class DispEntry(C.Structure):
    """mugiq_b200_disp_entry_t"""
    _fields_ = [("dir", C.c_int), ("sign", C.c_int), ("start", C.c_int), ("stop", C.c_int)]


class MugiqB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mugiq_b200 error {code}: {msg}")
        self.code = code


# every symbol include/mugiq_b200.h declares: (restype, argtypes)
_vp, _i, _ll, _d = C.c_void_p, C.c_int, C.c_longlong, C.c_double
_pg = C.POINTER(Geom)
_pe = C.POINTER(DispEntry)
_pvp = C.POINTER(C.c_void_p)
_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int)
SYMBOLS = {
    "mugiq_b200_version": (_i, []),
    "mugiq_b200_last_error": (C.c_char_p, []),
    "mugiq_b200_device_info": (_i, [C.c_char_p, _i, _pi, _pi, C.POINTER(_ll), C.POINTER(_ll)]),
    "mugiq_b200_gamma_tables": (_i, [_pd, _pi, _pd, _pi]),
    "mugiq_b200_ingest_spinor": (_i, [_vp, _vp, _i, _pg, _vp]),
    "mugiq_b200_export_spinor": (_i, [_vp, _i, _vp, _pg, _vp]),
    "mugiq_b200_ingest_spinor_batch": (_i, [_pvp, _pvp, _i, _i, _pg, _vp]),
    "mugiq_b200_gauge_upload": (_i, [_vp, _pvp, _pg, _vp]),
    "mugiq_b200_contract": (_i, [_vp, _vp, _vp, _d, _pg, _vp]),
    "mugiq_b200_contract_batch": (_i, [_vp, _pvp, _pvp, _pd, _i, _i, _pg, _vp]),
    "mugiq_b200_displace": (_i, [_vp, _vp, _vp, _i, _i, _pg, _vp]),
    "mugiq_b200_displace_batch": (_i, [_pvp, _pvp, _i, _vp, _i, _i, _pg, _vp]),
    "mugiq_b200_loop_workspace_bytes": (_ll, [_pg, _i, _pe, _i]),
    "mugiq_b200_loop_accumulate": (_i, [_vp, _pvp, _pd, _i, _vp, _pe, _i, _i, _vp, _pg, _vp]),
    "mugiq_b200_loop_plan_create": (_i, [C.POINTER(_vp), _vp, _pe, _i, _pg, _vp]),
    "mugiq_b200_loop_plan_destroy": (_i, [_vp]),
    "mugiq_b200_loop_plan_nloop": (_i, [_vp]),
    "mugiq_b200_loop_plan_info": (_i, [_vp, _pi, _pi, _pi, C.POINTER(_ll)]),
    "mugiq_b200_contract_native": (_i, [_vp, _pvp, _pvp, _pd, _i, _i, _i, _pg, _vp]),
    "mugiq_b200_displace_native": (_i, [_pvp, _pvp, _i, _vp, _i, _i, _i, _pg, _vp]),
    "mugiq_b200_peer_alloc": (_i, [C.POINTER(_vp), _ll, _vp]),
    "mugiq_b200_peer_open": (_i, [C.POINTER(_vp), _vp]),
    "mugiq_b200_peer_close": (_i, [_vp]),
    "mugiq_b200_peer_free": (_i, [_vp]),
    "mugiq_b200_halo_push_t": (_i, [_vp, _vp, _i, _i, _i, _ll, _i, _i, _i, _i, _i, _vp]),
    "mugiq_b200_fused_tiling_check": (_i, [_pe, _i, _pg, _i, _i, _i, _i, C.POINTER(_ll)]),
    "mugiq_b200_loop_plan_set_evec_order": (_i, [_vp, _i]),
    "mugiq_b200_loop_plan_computed_slots": (_i, [_vp, _pi, _i]),
    "mugiq_b200_loop_feed_create": (_i, [C.POINTER(_vp), _vp, _vp, _i, _i, _i, _i, _vp]),
    "mugiq_b200_loop_feed_destroy": (_i, [_vp]),
    "mugiq_b200_loop_feed_set_plan": (_i, [_vp, _vp, _vp]),
    "mugiq_b200_loop_feed_acquire": (_i, [_vp, _pvp, _i, _vp]),
    "mugiq_b200_loop_feed_commit": (_i, [_vp, _pd, _i, _vp]),
    "mugiq_b200_loop_feed_push_host": (_i, [_vp, _pvp, _pd, _i]),
    "mugiq_b200_loop_feed_finish": (_i, [_vp, C.POINTER(_ll)]),
    "mugiq_b200_comm_unique_id": (_i, [_vp]),
    "mugiq_b200_comm_create": (_i, [C.POINTER(_vp), _vp, _i, _i]),
    "mugiq_b200_comm_stage_bytes": (_ll, [_vp, _i, _i]),
    "mugiq_b200_comm_attach_peers": (_i, [_vp, _pvp, _pvp, _ll]),
    "mugiq_b200_comm_destroy": (_i, [_vp]),
    "mugiq_b200_comm_info": (_i, [_vp, _pi, _pi, _pi]),
    "mugiq_b200_allreduce": (_i, [_vp, _ll, _i, _vp, _vp]),
    "mugiq_b200_allgather": (_i, [_vp, _vp, _ll, _vp, _vp]),
    "mugiq_b200_allreduce_pos": (_i, [_vp, _pi, _i, _i, _i, _pg, _vp, _vp]),
    "mugiq_b200_loop_plan_accumulate_allreduce": (_i, [_vp, _vp, _pvp, _pd, _i, _i, _vp, _i, _vp]),
    "mugiq_b200_loop_plan_set_t_range": (_i, [_vp, _i, _i]),
    "mugiq_b200_loop_plan_t_halo": (_i, [_vp, _pi, _pi, _pi]),
    "mugiq_b200_loop_plan_accumulate": (_i, [_vp, _vp, _pvp, _pd, _i, _i, _vp]),
    "mugiq_b200_loop_plan_finalize": (_i, [_vp, _vp, _i, _vp]),
    "mugiq_b200_reorder_mapgamma": (_i, [_vp, _vp, _i, _i, _pg, _vp]),
    "mugiq_b200_phase_matrix": (_i, [_vp, _pi, _i, _i, _pi, _pi, _pi, _i, _vp]),
    "mugiq_b200_momproj_workspace_bytes": (_ll, [_ll, _i, _ll, _i]),
    "mugiq_b200_momproj": (_i, [_vp, _vp, _vp, _ll, _i, _ll, _i, _vp, _vp]),
    "mugiq_b200_phase_matrix_eo": (_i, [_vp, _pi, _i, _i, _pi, _pi, _pi, _i, _vp]),
    "mugiq_b200_momproj_pos_workspace_bytes": (_ll, [_pg, _i, _i]),
    "mugiq_b200_momproj_pos": (_i, [_vp, _vp, _vp, _i, _i, _pg, _vp, _vp]),
    "mugiq_b200_prof_enable": (_i, [_i]),
    "mugiq_b200_prof_reset": (_i, []),
    "mugiq_b200_prof_num_kernels": (_i, []),
    "mugiq_b200_prof_name": (C.c_char_p, [_i]),
    "mugiq_b200_prof_query": (_i, [_i, C.POINTER(_ll), C.POINTER(_ll), _pd, _pd, _pd]),
    "mugiq_b200_prof_fused_trace": (_i, [_vp, _ll]),
}

_lib = None


def load():
    """Load libmugiq_b200.so (once) and attach prototypes.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()' "
            "or make -C mugiq_b200/csrc).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    """Translate a negative status into an exception carrying mugiq_b200_last_error()."""
    if rc < 0:
        raise MugiqB200Error(rc, load().mugiq_b200_last_error().decode())
    return rc


def make_geom(L, precision=PREC_DOUBLE):
    g = Geom()
    for i in range(4):
        g.L[i] = int(L[i])
    g.precision = int(precision)
    return g


def ptr_array(ptrs):
    """HOST array of device pointers, as the batched entry points take."""
    arr = (C.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


def entry_array(entries):
    arr = (DispEntry * max(len(entries), 1))()
    for i, (d, s, a, b) in enumerate(entries):
        arr[i].dir, arr[i].sign, arr[i].start, arr[i].stop = int(d), int(s), int(a), int(b)
    return arr
