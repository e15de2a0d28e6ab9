"""The CPU oracle against golden vectors produced by the reference's OWN CUDA kernels (tests/golden/make_ref_golden.py,
run on a B200; oracle/_ref = the reference's lib/*.cu compiled unmodified against oracle/quda_shim).  This is what pins
the oracle: kernel arithmetic, gamma tables, neighbour / link / dagger selection, reorder map and phase formula are the
reference's; QUDA's accessor conventions are restated in the shim (quda_shim_core.h)."""
import os

import numpy as np
import pytest

from conftest import rel_err
from mugiq_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_kernels_4x4x4x8.npz")


@pytest.fixture(scope="module")
def gold():
    z = np.load(GOLD)
    L = tuple(int(x) for x in z["L"])
    ev = synth.random_evecs_np(L, int(z["nEv"]), seed=int(z["seed"]))
    U = synth.random_gauge(L, seed=int(z["seed"]))
    assert abs(ev.sum() - z["ev_checksum"]) < 1e-12 and abs(U.sum() - z["gauge_checksum"]) < 1e-9  # generator contract
    return z, L, ev, synth.sigmas(int(z["nEv"])), U


def test_oracle_loop_nest_matches_reference_kernels(oracle, gold):
    z, L, ev, sig, U = gold
    entries = [tuple(int(v) for v in e) for e in z["entries"]]
    out = oracle.compute_loop(ev, sig, U, entries, L)
    assert rel_err(out[:, :, ::5], z["dataPos_sample"]) < 1e-12
    assert rel_err(out.sum(axis=2), z["dataPos_sums"]) < 1e-12


def test_oracle_displacement_matches_reference_kernel(oracle, gold):
    z, L, ev, sig, U = gold
    for d in range(4):
        for s in (0, 1):
            got = oracle.displace(ev[0], U, d, s, L)
            assert rel_err(got[::3], z["displace_sample"][d, s]) < 1e-14, (d, s)
            assert rel_err(got.sum(axis=0), z["displace_sums"][d, s]) < 1e-13, (d, s)


def test_oracle_reorder_and_projection_match_reference_kernels(oracle, gold):
    z, L, ev, sig, U = gold
    entries = [tuple(int(v) for v in e) for e in z["entries"]]
    pos = oracle.compute_loop(ev, sig, U, entries, L)
    nLoop = pos.shape[0]
    mp = oracle.reorder_mapgamma(pos, nLoop, L).reshape(-1)
    assert rel_err(mp[::101], z["reorder_sample"]) < 1e-12
    w = (mp * np.arange(1, mp.size + 1)).sum()
    assert abs(w - z["reorder_weighted_sum"]) / abs(z["reorder_weighted_sum"]) < 1e-11  # position-sensitive checksum
    mom = z["mom"]
    for sgn, key in ((-1, "phase_minus"), (1, "phase_plus")):
        assert np.abs(oracle.phase_matrix(mom, sgn, L) - z[key]).max() < 5e-15
    M, N, K = L[3] * 16 * nLoop, len(mom), L[0] * L[1] * L[2]
    dm = oracle.gemm(oracle.reorder_mapgamma(pos, nLoop, L), oracle.phase_matrix(mom, -1, L), M, N, K)
    assert rel_err(dm.reshape(N, 16 * nLoop, L[3]), z["dataMom"]) < 1e-12
