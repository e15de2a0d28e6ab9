"""Schedule logic of the fused loop path (mugiq_b200_loop_plan_*), checked without a GPU: for entry lists that need no
Wilson-line storage (one-link plus hops read the gauge field in place) creating a plan launches nothing, so the
computed / derived / group bookkeeping can be inspected on the CPU."""
import ctypes as C

import pytest

from mugiq_b200 import _lib

FAKE_GAUGE = C.c_void_p(0x10000)  # never dereferenced: no kernel is launched for these plans


def plan_info(L, entries, gauge=FAKE_GAUGE):
    lib = _lib.load()
    h = C.c_void_p()
    g = _lib.make_geom(L)
    rc = lib.mugiq_b200_loop_plan_create(C.byref(h), gauge, _lib.entry_array(entries), len(entries), C.byref(g), None)
    if rc < 0:
        raise _lib.MugiqB200Error(rc, lib.mugiq_b200_last_error().decode())
    a, b, c, w = C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
    assert lib.mugiq_b200_loop_plan_info(h, C.byref(a), C.byref(b), C.byref(c), C.byref(w)) == 0
    n = lib.mugiq_b200_loop_plan_nloop(h)
    lib.mugiq_b200_loop_plan_destroy(h)
    return {"nloop": n, "computed": a.value, "derived": b.value, "groups": c.value, "wilson_bytes": w.value}


ONEHOP8 = [(d, s, 1, 1) for d in range(4) for s in (1, 0)]


def test_onehop8_uses_the_minus_from_plus_identity():
    info = plan_info((16, 16, 16, 32), ONEHOP8)
    # ultra-local + 4 plus loops computed in ONE launch group, the 4 minus loops derived after the eigenvector sum
    assert info == {"nloop": 9, "computed": 5, "derived": 4, "groups": 1, "wilson_bytes": 0}


def test_workspace_of_the_one_shot_call_covers_the_direct_minus_loops():
    """mugiq_b200_loop_accumulate(accumulate != 0) cannot derive minus loops from accumulated plus loops and computes
    them: each needs its daggered, shifted link field (144 B per site); plus one-link loops read the links in place."""
    lib = _lib.load()
    g = _lib.make_geom((16, 16, 16, 32))
    V4 = 16 * 16 * 16 * 32
    assert lib.mugiq_b200_loop_workspace_bytes(C.byref(g), 8, _lib.entry_array(ONEHOP8), 8) == 4 * V4 * 144
    plus = [(d, 1, 1, 1) for d in range(4)]
    assert lib.mugiq_b200_loop_workspace_bytes(C.byref(g), 8, _lib.entry_array(plus), 4) == 0
    # lengths 1..3 in +z: Wilson lines of 2 and 3 links
    assert lib.mugiq_b200_loop_workspace_bytes(C.byref(g), 8, _lib.entry_array([(2, 1, 1, 3)]), 1) == 2 * V4 * 144


def test_symmetry_switch_without_gpu_reports_the_allocation_failure(monkeypatch):
    monkeypatch.setenv("MUGIQ_B200_NO_PM_SYMMETRY", "1")
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    with pytest.raises(_lib.MugiqB200Error, match="Wilson-line storage"):
        plan_info((16, 16, 16, 32), ONEHOP8)


def test_repeated_entries_are_copies_and_ultralocal_only_is_one_group():
    assert plan_info((8, 8, 8, 8), [(0, 1, 1, 1), (0, 1, 1, 1), (2, 1, 1, 1)]) == {
        "nloop": 4, "computed": 3, "derived": 1, "groups": 1, "wilson_bytes": 0}
    assert plan_info((8, 8, 8, 8), [], gauge=None) == {"nloop": 1, "computed": 1, "derived": 0, "groups": 1, "wilson_bytes": 0}


def test_group_size_follows_the_lattice_row_length():
    # Lx = 48: a tile row pair fills the 8 warps with fewer loops per group than Lx = 16
    wide = plan_info((48, 4, 4, 4), [(d, 1, 1, 1) for d in range(4)])
    narrow = plan_info((16, 4, 4, 4), [(d, 1, 1, 1) for d in range(4)])
    assert narrow["groups"] == 1 and wide["groups"] >= narrow["groups"]
    assert wide["computed"] == narrow["computed"] == 5


@pytest.mark.parametrize("entries,msg", [([(4, 1, 1, 1)], "direction"), ([(0, 2, 1, 1)], "sign"), ([(0, 1, 3, 1)], "start")])
def test_bad_entries_are_rejected(entries, msg):
    with pytest.raises(_lib.MugiqB200Error, match=msg):
        plan_info((8, 8, 8, 8), entries)


def test_gauge_is_required_for_displacements():
    with pytest.raises(_lib.MugiqB200Error, match="gauge_d is NULL"):
        plan_info((8, 8, 8, 8), [(0, 1, 1, 1)], gauge=None)
