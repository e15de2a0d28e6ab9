"""Known-answer tests that pin the CPU oracle.  The reference ships no golden vectors,
so these algebraic identities — each derived from the reference source, not from its outputs — plus the
independent numpy restatement (test_oracle_vs_numpy.py) are what the oracle is pinned by."""
import numpy as np
import pytest

from conftest import rel_err
from mugiq_b200 import synth
from mugiq_b200.lattice import Lattice
from oracle import numpy_check as npc

L = (4, 4, 4, 8)


def dense_from_tables(rv, ci):
    g = np.zeros((16, 4, 4), dtype=complex)
    for n in range(16):
        for r in range(4):
            g[n, r, ci[n, r]] = rv[n, r, 0] + 1j * rv[n, r, 1]
    return g


def test_gamma_tables_match_product_definition(oracle):
    """(1) G(n) = g1^n0 g2^n1 g3^n2 g4^n3 in the DeGrand-Rossi basis (include/gamma.h:24-27)."""
    rv, ci, _, _ = oracle.gamma_tables()
    assert np.array_equal(dense_from_tables(rv, ci), npc.gamma_dense())


def test_clifford_algebra_and_gamma5(oracle):
    rv, ci, _, _ = oracle.gamma_tables()
    g = dense_from_tables(rv, ci)
    gs = [g[1], g[2], g[4], g[8]]
    for a in range(4):
        assert np.allclose(gs[a], gs[a].conj().T)  # hermitian
        for b in range(4):
            anti = gs[a] @ gs[b] + gs[b] @ gs[a]
            assert np.allclose(anti, 2 * np.eye(4) * (a == b))
    assert np.allclose(g[15], np.diag([1, 1, -1, -1]))  # g5 = g1 g2 g3 g4


def test_gamma_map_identity(oracle):
    """(2) g5 G(15-i) = sign[i] G(i) for every i except the two the reference documents as carrying the extra
    minus in the output name (i = 1, 4; include/gamma.h:90,93,97-98)."""
    rv, ci, sign, index = oracle.gamma_tables()
    g = dense_from_tables(rv, ci)
    assert list(index) == [15 - i for i in range(16)]
    assert [i for i in range(16) if sign[i] < 0] == [3, 6, 9, 11, 12, 14]
    for i in range(16):
        lhs = g[15] @ g[i]          # g5 * Gamma(i): the loop the output slot 15-i holds
        rhs = sign[i] * g[15 - i]
        if i in (1, 4):
            assert np.allclose(lhs, -rhs)
        else:
            assert np.allclose(lhs, rhs)


def test_coords_roundtrip(oracle):
    lat = Lattice((6, 4, 2, 4))
    c = lat.coords_eo()
    for xeo in range(0, lat.volume, 7):
        pty, cb = divmod(xeo, lat.volumeCB)
        x = oracle.get_coords(cb, lat.L, pty)
        assert list(c[xeo]) == x
        assert (sum(x) & 1) == pty
        assert oracle.cb_index(x, lat.L) == cb


def test_unit_gauge_is_pure_shift(oracle):
    """(3)"""
    lat = Lattice(L)
    v = synth.random_evecs_np(L, 1, seed=3)[0]
    U = synth.unit_gauge(L)
    for d in range(4):
        for s in (0, 1):
            w = oracle.displace(v, U, d, s, L)
            assert np.array_equal(w, v[lat.neighbour_eo(d, s)])


def test_minus_undoes_plus(oracle):
    """(4) D_{-mu} D_{+mu} v = v for unitary links."""
    v = synth.random_evecs_np(L, 1, seed=4)[0]
    U = synth.random_gauge(L, seed=4)
    for d in range(4):
        w = oracle.displace(oracle.displace(v, U, d, 1, L), U, d, 0, L)
        assert rel_err(w, v) < 1e-14
        w = oracle.displace(oracle.displace(v, U, d, 0, L), U, d, 1, L)
        assert rel_err(w, v) < 1e-14


def test_gauge_covariance(oracle):
    """(5) v -> Omega v, U_mu(x) -> Omega(x) U_mu(x) Omega(x+mu)^dag leaves every displaced loop invariant."""
    lat = Lattice(L)
    ev = synth.random_evecs_np(L, 3, seed=5)
    sig = synth.sigmas(3)
    U = synth.random_gauge(L, seed=5)
    Om = synth.random_gauge(L, seed=99)[0]  # [V4,3,3]
    ev2 = np.einsum("xrc,nxsc->nxsr", Om, ev.reshape(3, -1, 4, 3)).reshape(ev.shape)
    U2 = np.empty_like(U)
    for d in range(4):
        nb = lat.neighbour_eo(d, 1)
        U2[d] = np.einsum("xab,xbc,xdc->xad", Om, U[d], Om[nb].conj())
    entries = [(0, 1, 1, 2), (3, 0, 1, 1), (1, 0, 2, 3)]
    a = oracle.compute_loop(ev, sig, U, entries, L)
    b = oracle.compute_loop(np.ascontiguousarray(ev2), sig, np.ascontiguousarray(U2), entries, L)
    assert rel_err(b, a) < 1e-13


def test_ultralocal_reality_and_norm(oracle):
    """(6) hermitian Gamma -> real, anti-hermitian -> imaginary; sum_x T_1(x) = sum_n |v_n|^2 / sigma_n."""
    ev = synth.random_evecs_np(L, 4, seed=6)
    sig = synth.sigmas(4)
    out = oracle.compute_loop(ev, sig, None, [], L)[0]
    g = npc.gamma_dense()
    scale = np.abs(out).max()
    for G in range(16):
        if np.allclose(g[G], g[G].conj().T):
            assert np.abs(out[G].imag).max() < 1e-14 * scale
        else:
            assert np.allclose(g[G], -g[G].conj().T)
            assert np.abs(out[G].real).max() < 1e-14 * scale
    assert abs(out[0].sum() - (1.0 / sig).sum()) < 1e-10 * (1.0 / sig).sum()


def test_minus_loop_is_shifted_dagger_of_plus_loop(oracle):
    """Property used by the fused CUDA path: with M(x)[be,al] = v(x)^dag_be (D^k v)(x)_al,
    M_{-mu,k}(x) = M_{+mu,k}(x - k mu)^dagger.  Checked here on the gamma-projected loops through
    Gamma^dagger = +-Gamma."""
    lat = Lattice(L)
    ev = synth.random_evecs_np(L, 2, seed=8)
    sig = synth.sigmas(2)
    U = synth.random_gauge(L, seed=8)
    g = npc.gamma_dense()
    for d in range(4):
        for k in (1, 2):
            plus = oracle.compute_loop(ev, sig, U, [(d, 1, k, k)], L)[1]
            minus = oracle.compute_loop(ev, sig, U, [(d, 0, k, k)], L)[1]
            idx = np.arange(lat.volume)
            for _ in range(k):
                idx = lat.neighbour_eo(d, 0)[idx]  # x -> x - k mu
            for G in range(16):
                herm = 1.0 if np.allclose(g[G], g[G].conj().T) else -1.0
                # Tr[Gamma M^dag] = conj(Tr[Gamma^dag M]) = herm * conj(Tr[Gamma M])
                assert rel_err(minus[G], herm * plus[G][idx].conj()) < 1e-12


def test_momentum_projection_kats(oracle):
    """(7) p = 0 is the spatial sum; flipping FTSign conjugates the phases; the full momentum set on 4^3
    inverts exactly."""
    Ls = (4, 4, 4, 8)
    V3 = 64
    ph0 = oracle.phase_matrix([[0, 0, 0]], -1, Ls)
    assert np.array_equal(ph0, np.ones((1, V3)))
    moms = [[a, b, c] for a in range(4) for b in range(4) for c in range(4)]
    pm = oracle.phase_matrix(moms, -1, Ls)
    pp = oracle.phase_matrix(moms, +1, Ls)
    assert np.abs(pm - pp.conj()).max() < 1e-15
    rng = np.random.default_rng(0)
    M = 24
    A = rng.standard_normal((V3, M)) + 1j * rng.standard_normal((V3, M))  # memory m + M*k
    proj = oracle.gemm(A, pm, M, 64, V3)                                    # [N, M]
    back = oracle.gemm(np.ascontiguousarray(proj), np.ascontiguousarray(pp.T), M, V3, 64) / V3  # [V3, M]
    assert rel_err(back, A) < 1e-13
    assert rel_err(proj[0], A.sum(axis=0)) < 1e-14
