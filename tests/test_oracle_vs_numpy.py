"""The C++ oracle against the independent numpy restatement (oracle/numpy_check.py): same inputs, two
derivations that share neither code nor tables."""
import numpy as np
import pytest

from conftest import rel_err
from mugiq_b200 import synth
from oracle import numpy_check as npc

CASES = [((4, 4, 4, 8), 4), ((4, 2, 6, 4), 3), ((2, 2, 2, 2), 2), ((6, 4, 2, 4), 2)]


@pytest.mark.parametrize("L,nEv", CASES)
def test_compute_loop(oracle, L, nEv):
    ev = synth.random_evecs_np(L, nEv, seed=11)
    sig = synth.sigmas(nEv)
    U = synth.random_gauge(L, seed=11)
    entries = [(0, 1, 1, 1), (0, 0, 1, 1), (1, 1, 1, 2), (2, 0, 2, 2), (3, 1, 1, 1), (3, 0, 1, 3)]
    a = oracle.compute_loop(ev, sig, U, entries, L)
    b = npc.compute_loop(ev, sig, U, entries, L)
    assert a.shape == b.shape
    assert rel_err(a, b) < 1e-13


@pytest.mark.parametrize("L", [(4, 4, 4, 8), (4, 2, 6, 4)])
def test_displace_each_direction(oracle, L):
    v = synth.random_evecs_np(L, 1, seed=12)[0]
    U = synth.random_gauge(L, seed=12)
    for d in range(4):
        for s in (0, 1):
            assert rel_err(oracle.displace(v, U, d, s, L), npc.displace(v, U, d, s, L)) < 1e-14


@pytest.mark.parametrize("ftsign", [-1, 1])
def test_reorder_and_projection(oracle, ftsign):
    L = (4, 2, 6, 4)
    nLoop = 3
    rng = np.random.default_rng(5)
    V4 = L[0] * L[1] * L[2] * L[3]
    dataPos = rng.standard_normal((nLoop, 16, V4)) + 1j * rng.standard_normal((nLoop, 16, V4))
    mom = [[0, 0, 0], [1, 0, 0], [0, -1, 2], [1, 1, -1]]
    mp = oracle.reorder_mapgamma(dataPos, nLoop, L)
    ph = oracle.phase_matrix(mom, ftsign, L)
    M, N, K = L[3] * 16 * nLoop, len(mom), L[0] * L[1] * L[2]
    got = oracle.gemm(mp, ph, M, N, K).reshape(N, 16 * nLoop, L[3])
    ref = npc.momentum_projection(dataPos, mom, ftsign, L)
    assert rel_err(got, ref) < 1e-13


def test_float_oracle_tracks_double(oracle):
    L = (4, 4, 4, 4)
    ev = synth.random_evecs_np(L, 3, seed=13)
    sig = synth.sigmas(3)
    U = synth.random_gauge(L, seed=13)
    entries = [(2, 1, 1, 2)]
    a = oracle.compute_loop(ev, sig, U, entries, L)
    b = oracle.compute_loop(ev.astype(np.complex64), sig, U.astype(np.complex64), entries, L)
    assert b.dtype == np.complex64
    assert rel_err(b, a) < 1e-5
