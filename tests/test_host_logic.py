"""Host-side logic: displacement / momentum parsing, loop bookkeeping, lattice index helpers, sharding."""
import warnings

import numpy as np
import pytest

from mugiq_b200 import params as P
from mugiq_b200.dist import shard_range
from mugiq_b200.lattice import Lattice


def test_parse_disp_entries_reference_grammar():
    e, s, a, b = P.parse_disp_entries("+z:1,8;-x:3")
    assert e == ["+z:1,8", "-x:3"] and s == ["+z", "-x"] and a == [1, 3] and b == [8, 3]
    with pytest.raises(P.MugiqError):
        P.parse_disp_entries("")
    with pytest.raises(P.MugiqError):
        P.parse_disp_entries("+z")
    with pytest.raises(P.MugiqError):
        P.parse_disp_entries("+z:1,2,3")
    with pytest.raises(P.MugiqError):
        P.parse_disp_entries("+z:a")


def test_which_displace():
    assert [P.which_displace(s) for s in P.DISPLACE_FLAGS] == [(0, 1), (0, 0), (1, 1), (1, 0), (2, 1), (2, 0), (3, 1), (3, 0)]
    for bad in ("x", "+w", "++x", ""):
        with pytest.raises(P.MugiqError):
            P.which_displace(bad)


def test_loop_compute_param_counts_and_offsets():
    prm = P.MugiqLoopParam()
    prm.set_displacements("+z:1,8;-x:3;+t:2,4")
    c = P.LoopComputeParam(prm, (4, 4, 4, 8))
    assert c.nLoopPerEntry == [8, 1, 3] and c.nLoopOffset == [1, 9, 10] and c.nLoop == 13 and c.nData == 208
    assert c.entries() == [(2, 1, 1, 8), (0, 0, 3, 3), (3, 1, 2, 4)]
    assert c.loop_tags()[0] == "disp_0" and c.loop_tags()[8] == "disp_+z_8" and c.loop_tags()[-1] == "disp_+t_4"
    # ultra-local only
    c0 = P.LoopComputeParam(P.MugiqLoopParam(), (4, 4, 4, 8))
    assert c0.nLoop == 1 and c0.nDispEntries == 0 and c0.entries() == []


def test_start_stop_swap_warns():
    prm = P.MugiqLoopParam(doNonLocal=True, disp_entry=["+x:4,2"], disp_str=["+x"], disp_start=[4], disp_stop=[2])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        c = P.LoopComputeParam(prm, (4, 4, 4, 8))
    assert c.dispStart == [2] and c.dispStop == [4] and any("switch lengths" in str(x.message) for x in w)
    bad = P.MugiqLoopParam(doNonLocal=True, disp_str=["+x"], disp_start=[1, 2], disp_stop=[1])
    with pytest.raises(P.MugiqError):
        P.LoopComputeParam(bad, (4, 4, 4, 8))


def test_momenta_sets():
    assert [len(P.momenta_up_to(k)) for k in range(5)] == [1, 7, 19, 27, 33]
    assert P.momenta_up_to(0) == [[0, 0, 0]]


def test_read_momenta(tmp_path):
    f = tmp_path / "mom.txt"
    f.write_text("0 0 0\n1 0 -1\n")
    assert P.read_momenta(str(f)) == [[0, 0, 0], [1, 0, -1]]
    f.write_text("0 0\n")
    with pytest.raises(P.MugiqError):
        P.read_momenta(str(f))


def test_lattice_indexing():
    lat = Lattice((6, 4, 2, 4))
    c = lat.coords_eo()
    assert c.shape == (lat.volume, 4)
    assert np.array_equal(((c.sum(axis=1)) & 1), np.repeat([0, 1], lat.volumeCB))
    assert np.array_equal(lat.eo_index(c[:, 0], c[:, 1], c[:, 2], c[:, 3]), np.arange(lat.volume))
    assert np.array_equal(lat.eo_of_lex()[lat.lex_of_eo()], np.arange(lat.volume))
    for d in range(4):
        fwd, bwd = lat.neighbour_eo(d, 1), lat.neighbour_eo(d, 0)
        assert np.array_equal(bwd[fwd], np.arange(lat.volume))
    with pytest.raises(ValueError):
        Lattice((3, 4, 4, 4))


def test_shard_range_covers_everything():
    for nEv, world in [(16, 2), (200, 8), (7, 4), (3, 8), (1000, 8)]:
        spans = [shard_range(nEv, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == nEv
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
