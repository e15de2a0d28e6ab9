"""The oracle against the committed golden fixture of BASELINE.json configs[0] (4^3x8, 16 eigenvectors)."""
import os

import numpy as np

from conftest import rel_err
from mugiq_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "loop_4x4x4x8.npz")


def load_golden():
    z = np.load(GOLD)
    L = tuple(int(x) for x in z["L"])
    entries = [tuple(int(v) for v in e) for e in z["entries"]]
    ev = synth.random_evecs_np(L, int(z["nEv"]), seed=int(z["seed"]))
    U = synth.random_gauge(L, seed=int(z["seed"]))
    # the generator is part of the fixture contract
    assert abs(ev.sum() - z["ev_checksum"]) < 1e-12 and abs(U.sum() - z["gauge_checksum"]) < 1e-9
    return z, L, entries, ev, synth.sigmas(int(z["nEv"])), U


def test_oracle_reproduces_golden(oracle):
    z, L, entries, ev, sig, U = load_golden()
    out = oracle.compute_loop(ev, sig, U, entries, L)
    assert rel_err(out[:, :, ::37], z["dataPos_sample"]) < 1e-13
    assert rel_err(out.sum(axis=2), z["dataPos_sums"]) < 1e-12
    mom = z["mom"]
    mp = oracle.reorder_mapgamma(out, out.shape[0], L)
    ph = oracle.phase_matrix(mom, int(z["ftsign"]), L)
    M, N, K = L[3] * 16 * out.shape[0], len(mom), L[0] * L[1] * L[2]
    dm = oracle.gemm(mp, ph, M, N, K).reshape(N, 16 * out.shape[0], L[3])
    assert rel_err(dm, z["dataMom"]) < 1e-13
