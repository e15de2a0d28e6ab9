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


def test_group_size_does_not_depend_on_the_row_length():
    # a CTA works on a run of 32 consecutive checkerboard sites whatever Lx is: four loops per group everywhere
    wide = plan_info((48, 4, 4, 4), [(d, 1, 1, 1) for d in range(4)])
    narrow = plan_info((16, 4, 4, 4), [(d, 1, 1, 1) for d in range(4)])
    assert narrow["groups"] == 1 and wide["groups"] == 1
    assert wide["computed"] == narrow["computed"] == 5


def tiling(L, entries, t_begin=0, t_end=-1, precision=8, order=0):
    """mugiq_b200_fused_tiling_check for every launch group: [{run, units, nstages, stage_bytes, max_copies, mean_sites,
    misses, bad_maps}]."""
    lib = _lib.load()
    g = _lib.make_geom(L, precision)
    out = (C.c_longlong * 8)()
    ngroups = lib.mugiq_b200_fused_tiling_check(_lib.entry_array(entries), len(entries), C.byref(g), t_begin, t_end, -1, order, out)
    assert ngroups >= 1, lib.mugiq_b200_last_error()
    res = []
    for gi in range(ngroups):
        assert lib.mugiq_b200_fused_tiling_check(_lib.entry_array(entries), len(entries), C.byref(g), t_begin, t_end, gi, order,
                                                 out) == 0, lib.mugiq_b200_last_error()
        res.append(dict(zip(("run", "units", "nstages", "stage_bytes", "max_copies", "mean_sites", "misses", "bad_maps"), out)))
    return res


UP_TO_4 = [(d, s, 1, 4) for d in range(4) for s in (1, 0)]


@pytest.mark.parametrize("L,entries,t_range", [
    ((4, 4, 4, 8), ONEHOP8 + [(2, 1, 2, 3)], (0, -1)),     # BASELINE configs[0] (+ the smoke test's longer hops)
    ((16, 16, 16, 32), ONEHOP8, (0, -1)),                  # configs[1]
    ((24, 24, 24, 48), UP_TO_4, (0, -1)),                  # configs[2]: Lx/2 = 12, runs of 2 2/3 rows
    ((32, 32, 32, 64), [], (0, -1)),                       # configs[3]
    ((48, 48, 48, 16), ONEHOP8, (2, 14)),                  # configs[4]: extended time slab, interior only
    ((6, 6, 6, 6), UP_TO_4, (0, -1)),                      # Lx/2 = 3, volumeCB not a multiple of 32
    ((2, 4, 6, 4), UP_TO_4, (1, 3)),                       # one site per half-row
    ((12, 2, 2, 2), [(0, 1, 1, 7), (0, 0, 2, 9)], (0, -1)),  # x hops longer than the row
])
def test_every_thread_finds_its_sites_in_the_stage(L, entries, t_range):
    """Host replay of what thread 0 of every CTA builds (the merged intervals of a stage) and of what every thread looks
    up in it (own site, neighbour of every loop of the group): nothing may be missing, the maps must be sorted, disjoint,
    16-byte granular and fit the sized ring."""
    for prec in (8, 4):
        for t in tiling(L, entries, *t_range, precision=prec):
            assert t["misses"] == 0 and t["bad_maps"] == 0, t
            assert t["run"] == 16 * t["units"] and t["nstages"] >= 2 and t["max_copies"] <= 64, t


@pytest.mark.parametrize("L,entries,t_range", [
    ((4, 4, 4, 8), ONEHOP8 + [(2, 1, 2, 3)], (0, -1)),
    ((16, 16, 16, 32), ONEHOP8, (0, -1)),
    ((24, 24, 24, 48), UP_TO_4, (0, -1)),
    ((32, 32, 32, 64), [], (0, -1)),
    ((48, 48, 48, 16), ONEHOP8, (2, 14)),
    ((12, 2, 2, 2), [(0, 1, 1, 7), (0, 0, 2, 9)], (0, -1)),
])
def test_float2_stages_are_whole_chunks_fetched_as_tensor_boxes(L, entries, t_range):
    """Eigenvectors in QUDA FLOAT2 order: every interval of a stage is widened to chunks of 8 sites and fetched as tensor
    boxes of 4, 2 or 1 chunks; the boxes must tile the stage exactly and every thread's sites must still be in it."""
    for prec in (8, 4):
        for t in tiling(L, entries, *t_range, precision=prec, order=2):
            assert t["misses"] == 0 and t["bad_maps"] == 0, t
            assert t["nstages"] >= 2 and t["max_copies"] <= 64, t


def test_float2_staging_needs_whole_chunks_per_parity():
    lib = _lib.load()
    g = _lib.make_geom((6, 6, 6, 6), 8)   # volumeCB = 648 = 81 * 8: fine
    out = (C.c_longlong * 8)()
    assert lib.mugiq_b200_fused_tiling_check(_lib.entry_array(ONEHOP8), 8, C.byref(g), 0, -1, 0, 2, out) == 0
    # with even extents volumeCB is always a multiple of 8; odd extents (ultra-local loop only) can break it
    g = _lib.make_geom((2, 3, 3, 3), 8)   # volumeCB = 27
    assert lib.mugiq_b200_fused_tiling_check(_lib.entry_array([]), 0, C.byref(g), 0, -1, 0, 2, out) == -1
    assert b"multiple of 8" in lib.mugiq_b200_last_error()
    assert lib.mugiq_b200_fused_tiling_check(_lib.entry_array([]), 0, C.byref(g), 0, -1, 0, 0, out) == 0


def test_runs_fill_every_lane_on_the_baseline_lattices():
    """Groups of 4 loops: 32 sites per parity and CTA = 8 full warps, also where Lx/2 is 12 or 24 (whole-row tiles left
    8 of 32 lanes idle there); what a stage holds stays near 3.5x the run for mixed-direction groups."""
    for L, entries in [((16, 16, 16, 32), ONEHOP8), ((24, 24, 24, 48), UP_TO_4), ((48, 48, 48, 16), ONEHOP8)]:
        for t in tiling(L, entries):
            assert t["run"] == 32 and t["units"] == 2 and t["nstages"] >= 4, t
            assert t["mean_sites"] <= 2 * 32 * 4.2, t
    ul = tiling((32, 32, 32, 64), [])
    assert ul == [dict(run=128, units=8, nstages=4, stage_bytes=128 * 2 * 192, max_copies=2, mean_sites=256, misses=0,
                       bad_maps=0)]


def test_ultralocal_matrix_gets_its_own_role_beside_one_to_three_displaced_loops():
    """The first group carries the ultra-local loop.  With four displaced loops its matrix is shared out among their
    threads (4 roles x 2 warps); with one to three it has a role of its own, so the CTA's 8 compute warps are
    (loops + 1) roles x `units` warps - and every site of every role must still be in the stage."""
    L = (16, 16, 16, 32)
    for nd, units in ((1, 4), (2, 2), (3, 2), (4, 2)):
        t0 = tiling(L, [(d, 1, 1, 1) for d in range(nd)])[0]   # plus loops only: nothing is derived, nd computed + UL
        assert t0["units"] == units and t0["run"] == 16 * units and t0["misses"] == 0 and t0["bad_maps"] == 0, (nd, t0)
    # a second group never carries the ultra-local loop: 3 displaced loops keep 3 roles x 2 warps there too
    two_groups = tiling(L, [(d, 1, 1, 1) for d in range(4)] + [(0, 1, 2, 2), (1, 1, 2, 2), (2, 1, 2, 2)])
    assert len(two_groups) == 2 and all(t["misses"] == 0 and t["bad_maps"] == 0 for t in two_groups)


@pytest.mark.parametrize("entries,msg", [([(4, 1, 1, 1)], "direction"), ([(0, 2, 1, 1)], "sign"), ([(0, 1, 3, 1)], "start")])
def test_bad_entries_are_rejected(entries, msg):
    with pytest.raises(_lib.MugiqB200Error, match=msg):
        plan_info((8, 8, 8, 8), entries)


def test_gauge_is_required_for_displacements():
    with pytest.raises(_lib.MugiqB200Error, match="gauge_d is NULL"):
        plan_info((8, 8, 8, 8), [(0, 1, 1, 1)], gauge=None)
