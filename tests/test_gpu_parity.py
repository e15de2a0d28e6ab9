"""Parity of the CUDA path (through the C-ABI) against the CPU oracle on identical seeded inputs.
FP64 criterion: max|delta| / max|ref| <= 1e-12 (BASELINE.json north_star); FP32: 2e-5."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import rel_err, TOL_F64, TOL_F32
from mugiq_b200 import synth, _lib
from mugiq_b200.lattice import Lattice
from mugiq_b200.params import MugiqLoopParam, momenta_up_to, MugiqError

pytestmark = pytest.mark.gpu

LATTICES = [(4, 4, 4, 8), (4, 2, 6, 4), (2, 2, 2, 2), (8, 4, 4, 6)]


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from mugiq_b200 import ops as o
    info = o.device_info()
    assert info["cc"] >= 100, f"expected an sm_100 device, got {info}"
    return o


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


def cdt(prec):
    return np.complex128 if prec == 8 else np.complex64


@pytest.mark.parametrize("L", LATTICES)
@pytest.mark.parametrize("prec", [8, 4])
def test_contract_single_pair(ops, oracle, L, prec):
    """performLoopContraction: accumulating call, vL != vR and vL == vR."""
    ev = synth.random_evecs_np(L, 2, seed=31).astype(cdt(prec))
    V4 = ev.shape[1]
    rng = np.random.default_rng(1)
    init = (rng.standard_normal((16, V4)) + 1j * rng.standard_normal((16, V4))).astype(cdt(prec)) * 0.01
    tol = TOL_F64 if prec == 8 else TOL_F32
    for (a, b) in [(0, 1), (0, 0)]:
        ref = oracle.contract(init.copy(), ev[a], ev[b], 0.37, L)
        loop = dev(init)
        vl, vr = dev(ev[a]), (dev(ev[b]) if b != a else None)
        ops.contract(loop, vl, vr if vr is not None else vl, 0.37, L)
        assert rel_err(host(loop), ref) < tol


@pytest.mark.parametrize("L", LATTICES)
@pytest.mark.parametrize("prec", [8, 4])
def test_displace_all_directions(ops, oracle, L, prec):
    v = synth.random_evecs_np(L, 1, seed=32)[0].astype(cdt(prec))
    U = synth.random_gauge(L, seed=32).astype(cdt(prec))
    gd = ops.gauge_upload(U, L)
    vd = dev(v)
    tol = 1e-14 if prec == 8 else 1e-6
    for d in range(4):
        for s in (0, 1):
            out = torch.empty_like(vd)
            ops.displace(out, vd, gd, d, s, L)
            assert rel_err(host(out), oracle.displace(v, U, d, s, L)) < tol


@pytest.mark.parametrize("L", [(16, 16, 2, 2), (24, 6, 2, 2), (6, 4, 2, 4), (8, 2, 2, 2), (32, 4, 2, 2)])
@pytest.mark.parametrize("prec", [8, 4])
def test_displace_batch_tiles(ops, oracle, L, prec):
    """Batched hop on lattices whose y extent spans several tiles (wrap inside and across tiles), with odd Lx/2
    (per-site kernel in FP32) and with more eigenvectors than pipeline stages; dst buffers are pre-filled with
    garbage so that an unwritten site shows."""
    n = 5
    v = synth.random_evecs_np(L, n, seed=41).astype(cdt(prec))
    U = synth.random_gauge(L, seed=41).astype(cdt(prec))
    gd = ops.gauge_upload(U, L)
    vd = [dev(v[i]) for i in range(n)]
    tol = 1e-14 if prec == 8 else 1e-6
    for d in range(4):
        for s in (0, 1):
            outs = [torch.full_like(vd[0], 7.0) for _ in range(n)]
            ops.displace_batch(outs, vd, gd, d, s, L)
            for i in range(n):
                assert rel_err(host(outs[i]), oracle.displace(v[i], U, d, s, L)) < tol, (d, s, i)


@pytest.mark.parametrize("L", [(10, 6, 4, 4), (12, 5, 3, 4), (20, 7, 2, 2), (36, 3, 2, 2), (64, 2, 2, 2), (128, 2, 1, 2), (4, 16, 2, 2),
                               (6, 9, 2, 2), (2, 3, 5, 2), (20, 14, 2, 2), (36, 6, 2, 2), (2, 6, 10, 2), (128, 2, 2, 2)])
def test_stage_kernels_tile_geometries(ops, oracle, L):
    """Tile selection of the bulk-TMA stage kernels over awkward extents: odd and prime Ly (one row per tile, or the whole
    y range in one tile with the wrap inside it), half-rows longer than the tile target, Lx/2 = 1, 3, 5, 9, 32, 64."""
    n = 3
    v = synth.random_evecs_np(L, n, seed=45)
    U = synth.random_gauge(L, seed=45)
    gd = ops.gauge_upload(U, L)
    vd = [dev(v[i]) for i in range(n)]
    if all(x % 2 == 0 for x in L):
        for d in range(4):
            for s in (0, 1):
                outs = [torch.full_like(vd[0], 7.0) for _ in range(n)]
                ops.displace_batch(outs, vd, gd, d, s, L)
                for i in range(n):
                    assert rel_err(host(outs[i]), oracle.displace(v[i], U, d, s, L)) < 1e-14, (d, s, i)
    else:  # odd Ly / Lz / Lt: layout-only kernels below; displacements are rejected (QUDA needs even extents)
        with pytest.raises(_lib.MugiqB200Error, match="must be even for displacements"):
            ops.displace_batch([torch.empty_like(vd[0])], vd[:1], gd, 1, 1, L)
    V4 = v.shape[1]
    ref = np.zeros((16, V4), dtype=np.complex128)
    sig = [0.4, 0.5, 0.6]
    for i in range(n):
        ref = oracle.contract(ref, v[i], v[(i + 1) % n], sig[i], L)
    loop = torch.full((16, V4), 3.0, dtype=torch.complex128, device="cuda")
    ops.contract_batch(loop, vd, [vd[(i + 1) % n] for i in range(n)], sig, L, accumulate=False)
    assert rel_err(host(loop), ref) < TOL_F64
    nLoop = 2
    rng = np.random.default_rng(6)
    pos = rng.standard_normal((nLoop, 16, V4)) + 1j * rng.standard_normal((nLoop, 16, V4))
    out = torch.zeros(V4 * 16 * nLoop, dtype=torch.complex128, device="cuda")
    ops.reorder_mapgamma(out, dev(pos), 16 * nLoop, nLoop, L)
    assert np.array_equal(host(out), oracle.reorder_mapgamma(pos, nLoop, L).reshape(-1))
    mom = momenta_up_to(2)
    got = ops.momproj_pos(dev(pos), ops.phase_matrix_eo(mom, -1, L), nLoop, L)
    from oracle import numpy_check as npc
    assert rel_err(host(got), npc.momentum_projection(pos, mom, -1, L)) < TOL_F64


@pytest.mark.parametrize("L", [(6, 2, 2, 3), (2, 2, 2, 2), (16, 4, 4, 4)])
@pytest.mark.parametrize("same", [True, False])
def test_contract_batch_ragged_tiles(ops, oracle, L, same):
    """Batched contraction with more eigenvectors than shared-memory stages and a last tile that is not full."""
    n = 11
    ev = synth.random_evecs_np(L, 2 * n, seed=43)
    sig = [0.2 + 0.05 * i for i in range(n)]
    V4 = ev.shape[1]
    ref = np.zeros((16, V4), dtype=np.complex128)
    for i in range(n):
        ref = oracle.contract(ref, ev[i], ev[i] if same else ev[n + i], sig[i], L)
    loop = torch.full((16, V4), 3.0, dtype=torch.complex128, device="cuda")
    vl = [dev(ev[i]) for i in range(n)]
    vr = None if same else [dev(ev[n + i]) for i in range(n)]
    ops.contract_batch(loop, vl, vr, sig, L, accumulate=False)
    assert rel_err(host(loop), ref) < TOL_F64
    ops.contract_batch(loop, vl, vr, sig, L, accumulate=True)
    assert rel_err(host(loop), 2 * ref) < TOL_F64


def test_unaligned_fp32_fields_take_the_per_site_kernels(ops, oracle):
    """FP32 fields that are only 8-byte aligned cannot be fetched by bulk TMA (16-byte granularity): the per-site
    kernels serve them, with the same results."""
    L = (4, 4, 4, 8)
    ev = synth.random_evecs_np(L, 2, seed=44).astype(np.complex64)
    U = synth.random_gauge(L, seed=44).astype(np.complex64)
    V4 = ev.shape[1]
    buf = torch.zeros(3 * V4 * 12 + 1, dtype=torch.complex64, device="cuda")
    f = [buf[1 + k * V4 * 12:1 + (k + 1) * V4 * 12].view(V4, 12) for k in range(3)]
    assert all(t.data_ptr() % 16 == 8 for t in f)
    f[0].copy_(dev(ev[0]))
    f[1].copy_(dev(ev[1]))
    loop = torch.zeros((16, V4), dtype=torch.complex64, device="cuda")
    ops.contract(loop, f[0], f[1], 0.7, L)
    assert rel_err(host(loop), oracle.contract(np.zeros((16, V4), dtype=np.complex64), ev[0], ev[1], 0.7, L)) < TOL_F32
    gd = ops.gauge_upload(U, L)
    for d, s in ((0, 1), (3, 0)):
        ops.displace(f[2], f[0], gd, d, s, L)
        assert rel_err(host(f[2]), oracle.displace(ev[0], U, d, s, L)) < 1e-6


def test_contract_batch_accumulate_and_overwrite(ops, oracle):
    L = (4, 4, 4, 8)
    n = 7
    ev = synth.random_evecs_np(L, n, seed=33)
    ev2 = synth.random_evecs_np(L, n, seed=34)
    sig = synth.sigmas(n)
    V4 = ev.shape[1]
    ref = np.zeros((16, V4), dtype=np.complex128)
    for i in range(n):
        oracle.contract(ref, ev[i], ev2[i], sig[i], L)
    vl = [dev(ev[i]) for i in range(n)]
    vr = [dev(ev2[i]) for i in range(n)]
    loop = torch.full((16, V4), 7.0 + 1j, dtype=torch.complex128, device="cuda")
    ops.contract_batch(loop, vl, vr, sig, L, accumulate=False)
    assert rel_err(host(loop), ref) < TOL_F64
    ops.contract_batch(loop, vl, vr, sig, L, accumulate=True)
    assert rel_err(host(loop), 2 * ref) < TOL_F64
    # ultra-local form (vR = NULL)
    ref0 = np.zeros((16, V4), dtype=np.complex128)
    for i in range(n):
        oracle.contract(ref0, ev[i], ev[i], sig[i], L)
    ops.contract_batch(loop, vl, None, sig, L, accumulate=False)
    assert rel_err(host(loop), ref0) < TOL_F64
    # empty batch: overwrite -> zeros, accumulate -> untouched
    ops.contract_batch(loop, [], None, [], L, accumulate=True)
    assert rel_err(host(loop), ref0) < TOL_F64
    ops.contract_batch(loop, [], None, [], L, accumulate=False)
    assert float(loop.abs().max()) == 0.0


ENTRY_SETS = {
    "ultralocal": [],
    "onehop8": [(0, 1, 1, 1), (0, 0, 1, 1), (1, 1, 1, 1), (1, 0, 1, 1), (2, 1, 1, 1), (2, 0, 1, 1), (3, 1, 1, 1), (3, 0, 1, 1)],
    "ranges": [(2, 1, 1, 3), (0, 0, 2, 2), (3, 0, 1, 2), (1, 1, 2, 4)],
    "plus_only": [(0, 1, 1, 1), (3, 1, 1, 2)],
    "minus_only": [(1, 0, 1, 1), (2, 0, 1, 3)],
    "repeated": [(0, 1, 1, 1), (0, 1, 1, 1), (0, 0, 1, 2), (0, 1, 2, 2)],
}


@pytest.mark.parametrize("L", LATTICES)
@pytest.mark.parametrize("name", list(ENTRY_SETS))
def test_loop_accumulate_matches_oracle(ops, oracle, L, name):
    entries = ENTRY_SETS[name]
    nEv = 5
    ev = synth.random_evecs_np(L, nEv, seed=35)
    sig = synth.sigmas(nEv)
    U = synth.random_gauge(L, seed=35)
    ref = oracle.compute_loop(ev, sig, U, entries, L)
    gd = ops.gauge_upload(U, L)
    evd = [dev(ev[i]) for i in range(nEv)]
    out = torch.full(ref.shape, 3.0, dtype=torch.complex128, device="cuda")
    ops.loop_accumulate(out, evd, sig, gd, entries, L, accumulate=False)
    assert rel_err(host(out), ref) < TOL_F64
    # split into two batches, second accumulating: same sum
    out2 = torch.full(ref.shape, -1.0, dtype=torch.complex128, device="cuda")
    ops.loop_accumulate(out2, evd[:2], sig[:2], gd, entries, L, accumulate=False)
    ops.loop_accumulate(out2, evd[2:], sig[2:], gd, entries, L, accumulate=True)
    assert rel_err(host(out2), ref) < TOL_F64


LOOP_LATTICES = [(16, 4, 4, 4), (12, 2, 4, 2), (6, 4, 2, 4), (24, 2, 2, 2), (4, 4, 4, 8), (8, 2, 2, 6), (6, 6, 6, 2), (48, 2, 2, 2)]


@pytest.mark.parametrize("L", LOOP_LATTICES)
@pytest.mark.parametrize("name", ["onehop8", "ranges", "repeated"])
@pytest.mark.parametrize("symmetric", [True, False])
def test_loop_plan_batches(ops, oracle, L, name, symmetric, monkeypatch):
    """Plan form: Wilson lines built once, eigenvectors fed in three batches, derived slots filled by finalize;
    with and without the minus-from-plus identity; lattices exercising every lane->site mapping of the fused
    kernel (Lx/2 = 8, 6, 3, 12, 2, 4; tiles 2x2x1, 2x1x2, 1x2x2 ...)."""
    if not symmetric:
        monkeypatch.setenv("MUGIQ_B200_NO_PM_SYMMETRY", "1")
    entries = ENTRY_SETS[name]
    nEv = 7
    ev = synth.random_evecs_np(L, nEv, seed=41)
    sig = synth.sigmas(nEv)
    U = synth.random_gauge(L, seed=41)
    ref = oracle.compute_loop(ev, sig, U, entries, L)
    gd = ops.gauge_upload(U, L)
    evd = [dev(ev[i]) for i in range(nEv)]
    plan = ops.LoopPlan(gd, entries, L)
    info = plan.info()
    assert plan.nLoop == ref.shape[0]
    if symmetric and name == "onehop8":
        assert info["computed"] == 5 and info["derived"] == 4 and info["wilson_bytes"] == 0
    if not symmetric and name == "onehop8":
        assert info["computed"] == 9 and info["derived"] == 0
    out = torch.full(ref.shape, 5.0 - 2.0j, dtype=torch.complex128, device="cuda")
    plan.accumulate(out, evd[:3], sig[:3], accumulate=False)
    plan.accumulate(out, evd[3:4], sig[3:4], accumulate=True)
    plan.accumulate(out, evd[4:], sig[4:], accumulate=True)
    plan.finalize(out)
    assert rel_err(host(out), ref) < TOL_F64
    plan.close()


def test_loop_plan_many_derived_loops(ops, oracle):
    """More minus loops derived from their plus partners than one finalize launch takes (kMinusBatch = 32): 36 + 3 pairs."""
    L = (4, 4, 8, 40)
    nEv = 2
    ev = synth.random_evecs_np(L, nEv, seed=43)
    sig = synth.sigmas(nEv)
    U = synth.random_gauge(L, seed=43)
    entries = [(3, 1, 1, 36), (3, 0, 1, 36), (2, 1, 1, 3), (2, 0, 1, 3)]
    ref = oracle.compute_loop(ev, sig, U, entries, L)
    gd = ops.gauge_upload(U, L)
    plan = ops.LoopPlan(gd, entries, L)
    assert plan.info()["derived"] == 39
    out = torch.full(ref.shape, 1.0 + 1.0j, dtype=torch.complex128, device="cuda")
    plan.accumulate(out, [dev(ev[i]) for i in range(nEv)], sig, accumulate=False)
    plan.finalize(out)
    assert rel_err(host(out), ref) < TOL_F64
    plan.close()


def test_fused_trace_timeline(ops):
    """mugiq_b200_prof_fused_trace: every CTA of a fused launch stamps its SM and five ordered times; NULL switches it off."""
    L = (8, 8, 8, 8)
    nEv = 6
    ev = dev(synth.random_evecs_np(L, nEv, seed=44))
    sig = synth.sigmas(nEv)
    gd = ops.gauge_upload(synth.random_gauge(L, seed=44), L)
    plan = ops.LoopPlan(gd, [(0, 1, 1, 1), (1, 1, 1, 1), (2, 1, 1, 1), (3, 1, 1, 1)], L)
    out = torch.zeros((plan.nLoop, 16, ev.shape[1]), dtype=torch.complex128, device="cuda")
    ncta = 8 * 8 * 8 * 8 // 2 // 32
    buf = torch.zeros(16 * (ncta + 4), dtype=torch.int64, device="cuda")
    ops.prof_fused_trace(buf)
    plan.accumulate(out, list(ev), sig, accumulate=False)
    torch.cuda.synchronize()
    ops.prof_fused_trace(None)
    t = buf.cpu().numpy().reshape(-1, 16)
    assert (t[:ncta, 1] > 0).all() and (t[ncta:] == 0).all()
    assert (t[:ncta, 0] >= 0).all() and (t[:ncta, 0] < 256).all()
    for k in range(1, 5):
        assert (t[:ncta, k + 1] >= t[:ncta, k]).all() and (t[:ncta, 9 + k] > t[:ncta, 8 + k]).all()
    buf.zero_()
    plan.accumulate(out, list(ev), sig, accumulate=False)
    torch.cuda.synchronize()
    assert int(buf.abs().sum().item()) == 0
    plan.close()


def test_loop_plan_many_vectors(ops, oracle):
    """More eigenvectors than one launch's pointer table (kFusedMaxVec = 256): chunked launches accumulate."""
    L = (4, 2, 2, 2)
    nEv = 300
    ev = synth.random_evecs_np(L, nEv, seed=42)
    sig = synth.sigmas(nEv)
    U = synth.random_gauge(L, seed=42)
    entries = [(1, 1, 1, 2), (1, 0, 2, 2)]
    ref = oracle.compute_loop(ev, sig, U, entries, L)
    gd = ops.gauge_upload(U, L)
    evd = dev(ev)
    out = torch.zeros(ref.shape, dtype=torch.complex128, device="cuda")
    ops.loop_accumulate(out, list(evd), sig, gd, entries, L)
    assert rel_err(host(out), ref) < TOL_F64


def test_loop_accumulate_float(ops, oracle):
    L = (4, 4, 4, 8)
    entries = ENTRY_SETS["onehop8"] + [(3, 1, 2, 3)]
    ev = synth.random_evecs_np(L, 6, seed=36)
    sig = synth.sigmas(6)
    U = synth.random_gauge(L, seed=36)
    ref = oracle.compute_loop(ev, sig, U, entries, L)
    gd = ops.gauge_upload(U.astype(np.complex64), L)
    evd = [dev(ev[i].astype(np.complex64)) for i in range(6)]
    out = torch.zeros(ref.shape, dtype=torch.complex64, device="cuda")
    ops.loop_accumulate(out, evd, sig, gd, entries, L)
    assert rel_err(host(out), ref) < TOL_F32


def test_golden_fixture_16_evecs(ops):
    """BASELINE.json configs[0] (4^3x8, 16 eigenvectors) against the committed golden vectors."""
    from test_golden import load_golden
    z, L, entries, ev, sig, U = load_golden()
    gd = ops.gauge_upload(U, L)
    evd = [dev(ev[i]) for i in range(len(ev))]
    nLoop = 1 + sum(b - a + 1 for (_, _, a, b) in entries)
    out = torch.zeros((nLoop, 16, ev.shape[1]), dtype=torch.complex128, device="cuda")
    ops.loop_accumulate(out, evd, sig, gd, entries, L)
    h = host(out)
    assert rel_err(h[:, :, ::37], z["dataPos_sample"]) < TOL_F64
    assert rel_err(h.sum(axis=2), z["dataPos_sums"]) < TOL_F64
    mp = torch.empty((L[0] * L[1] * L[2], 16 * nLoop, L[3]), dtype=torch.complex128, device="cuda")
    ops.reorder_mapgamma(mp, out, 16 * nLoop, nLoop, L)
    ph = ops.phase_matrix(z["mom"], int(z["ftsign"]), L)
    dm = ops.momproj(mp, ph, L[3] * 16 * nLoop, len(z["mom"]), L[0] * L[1] * L[2])
    assert rel_err(host(dm).reshape(z["dataMom"].shape), z["dataMom"]) < TOL_F64


@pytest.mark.parametrize("L", [(4, 4, 4, 8), (4, 2, 6, 4), (8, 4, 4, 6)])
@pytest.mark.parametrize("prec", [8, 4])
def test_reorder_mapgamma(ops, oracle, L, prec):
    nLoop = 3
    V4 = Lattice(L).volume
    rng = np.random.default_rng(2)
    inp = (rng.standard_normal((nLoop, 16, V4)) + 1j * rng.standard_normal((nLoop, 16, V4))).astype(cdt(prec))
    ref = oracle.reorder_mapgamma(inp, nLoop, L)
    out = torch.zeros(ref.shape, dtype=torch.complex128 if prec == 8 else torch.complex64, device="cuda")
    ops.reorder_mapgamma(out, dev(inp), 16 * nLoop, nLoop, L)
    assert np.array_equal(host(out), ref)  # pure data movement and sign flips: bit-exact
    with pytest.raises(_lib.MugiqB200Error, match="nData = nLoop"):
        ops.reorder_mapgamma(out, dev(inp), 16 * nLoop + 1, nLoop, L)


@pytest.mark.parametrize("prec", [8, 4])
@pytest.mark.parametrize("ftsign", [-1, 1])
def test_phase_matrix(ops, oracle, prec, ftsign):
    L = (4, 6, 8, 4)
    tot = (8, 6, 16, 4)
    cc = (1, 0, 1, 0)
    mom = momenta_up_to(4)
    ref = oracle.phase_matrix(mom, ftsign, L, tot, cc, dtype=cdt(prec))
    out = ops.phase_matrix(mom, ftsign, L, tot, cc, dtype=torch.complex128 if prec == 8 else torch.complex64)
    # FP32: the reference (and the oracle) accumulate the phase in float, this kernel in double before rounding
    assert np.abs(host(out) - ref).max() < (4e-15 if prec == 8 else 5e-6)


@pytest.mark.parametrize("M,N,K", [(256, 1, 64), (4608, 1, 4096), (96, 7, 128), (1000, 19, 333), (512, 33, 512),
                                   (130, 40, 70), (8, 3, 5)])
def test_momproj_f64(ops, oracle, M, N, K):
    rng = np.random.default_rng(3)
    A = rng.standard_normal((K, M)) + 1j * rng.standard_normal((K, M))   # memory m + M*k
    B = rng.standard_normal((N, K)) + 1j * rng.standard_normal((N, K))   # memory k + K*n
    ref = oracle.gemm(A, B, M, N, K)
    out = ops.momproj(dev(A), dev(B), M, N, K)
    assert rel_err(host(out), ref) < TOL_F64


@pytest.mark.parametrize("L", [(4, 4, 4, 8), (8, 2, 6, 4), (6, 4, 2, 3), (16, 4, 4, 4)])
@pytest.mark.parametrize("nmom,prec", [(1, 8), (7, 8), (33, 8), (40, 8), (7, 4)])
def test_momproj_pos_fused(ops, oracle, L, nmom, prec):
    """Stages 3+4 in one kernel (momentum projection straight from dataPos) against the oracle's reorder + GEMM."""
    nLoop = 3
    lat = Lattice(L)
    rng = np.random.default_rng(5)
    pos = (rng.standard_normal((nLoop, 16, lat.volume)) + 1j * rng.standard_normal((nLoop, 16, lat.volume))).astype(cdt(prec))
    mom = rng.integers(-3, 4, size=(nmom, 3)).tolist()
    mp = oracle.reorder_mapgamma(pos.astype(np.complex128), nLoop, L)
    ph = oracle.phase_matrix(mom, -1, L)
    ref = oracle.gemm(mp, ph, L[3] * 16 * nLoop, nmom, lat.V3).reshape(nmom, 16 * nLoop, L[3])
    ph_eo = ops.phase_matrix_eo(mom, -1, L, dtype=torch.complex128 if prec == 8 else torch.complex64)
    out = ops.momproj_pos(dev(pos), ph_eo, nLoop, L)
    assert rel_err(host(out), ref) < (TOL_F64 if prec == 8 else TOL_F32)


def test_momproj_f32(ops, oracle):
    M, N, K = 384, 7, 256
    rng = np.random.default_rng(4)
    A = (rng.standard_normal((K, M)) + 1j * rng.standard_normal((K, M))).astype(np.complex64)
    B = (rng.standard_normal((N, K)) + 1j * rng.standard_normal((N, K))).astype(np.complex64)
    ref = oracle.gemm(A.astype(np.complex128), B.astype(np.complex128), M, N, K)
    out = ops.momproj(dev(A), dev(B), M, N, K)
    assert rel_err(host(out), ref) < TOL_F32


@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("prec", [8, 4])
def test_ingest_export_quda_orders(ops, order, prec):
    L = (4, 2, 6, 4)
    lat = Lattice(L)
    v = synth.random_evecs_np(L, 1, seed=37)[0].astype(cdt(prec))  # site-major [V4,12]
    s = v.reshape(2, lat.volumeCB, 12)
    if order == 2:   # [parity][comp][x_cb]
        q = np.ascontiguousarray(s.transpose(0, 2, 1))
    else:            # [parity][j][x_cb][2]
        q = np.ascontiguousarray(s.reshape(2, lat.volumeCB, 6, 2).transpose(0, 2, 1, 3))
    got = ops.ingest_spinor(dev(q.reshape(-1, 12)), order, L)
    assert np.array_equal(host(got), v)
    back = ops.export_spinor(got, order, L)
    assert np.array_equal(host(back).ravel(), q.ravel())


def test_error_reporting(ops):
    L = (4, 4, 4, 4)
    V4 = 256
    v = torch.zeros((V4, 12), dtype=torch.complex128, device="cuda")
    g = torch.zeros((4, V4, 3, 3), dtype=torch.complex128, device="cuda")
    with pytest.raises(_lib.MugiqB200Error, match="direction"):
        ops.displace(torch.empty_like(v), v, g, 4, 1, L)
    with pytest.raises(_lib.MugiqB200Error, match="differ"):
        ops.displace(v, v, g, 0, 1, L)
    with pytest.raises(_lib.MugiqB200Error, match="even"):
        ops.displace(torch.empty_like(v), v, g, 0, 1, (3, 4, 4, 4))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.displace(torch.empty_like(v), v.cpu(), g, 0, 1, L)


def test_loop_mugiq_end_to_end(ops, oracle, tmp_path):
    """The Loop_Mugiq mirror: device-resident and host-streamed eigenvectors, momentum projection, output."""
    from mugiq_b200.loop import Loop_Mugiq, Eigsolve, computeLoop
    from oracle import numpy_check as npc
    L = (4, 4, 4, 8)
    nEv = 9
    ev = synth.random_evecs_np(L, nEv, seed=38)
    sig = synth.sigmas(nEv)
    U = synth.random_gauge(L, seed=38)
    prm = MugiqLoopParam(gauge=[U[mu] for mu in range(4)])
    prm.set_displacements("+x:1;-y:1,2;+t:2")
    mom = momenta_up_to(1)
    prm.set_momenta(mom)
    prm.writeMomSpaceHDF5 = True
    prm.fname_mom_h5 = str(tmp_path / "loops.npz")
    entries = [(0, 1, 1, 1), (1, 0, 1, 2), (3, 1, 2, 2)]
    ref = oracle.compute_loop(ev, sig, U, entries, L)
    ref_mom = npc.momentum_projection(ref, mom, -1, L)
    for resident, fused in ((True, True), (False, True), (True, False)):
        vecs = [dev(ev[i]) if resident else torch.from_numpy(ev[i]).pin_memory() for i in range(nEv)]
        loop = computeLoop(prm, Eigsolve(vecs, sig, L), evec_batch=4, fused_momproj=fused)
        assert rel_err(loop.dataPos.numpy(), ref) < TOL_F64
        assert rel_err(loop.dataMom.numpy(), ref_mom) < TOL_F64
        with pytest.raises(MugiqError, match="more than once"):
            loop.performMomentumProjection()
    z = np.load(prm.fname_mom_h5)
    key = "mom_+1_+0_+0/disp_-y_2/g5g4/loop"
    im = mom.index([1, 0, 0])
    assert np.allclose(z[key][:, 0] + 1j * z[key][:, 1], ref_mom[im, 7 + 16 * 3, :], rtol=0, atol=1e-12 * np.abs(ref_mom).max())
    # the same tree as a real HDF5 file (self-contained writer): every dataset of writeLoopsHDF5_Mom, [T][2] doubles
    from mugiq_b200 import h5min
    prm.fname_mom_h5 = str(tmp_path / "loops.h5")
    loop.momSpaceFilename = prm.fname_mom_h5
    loop.writeLoopsHDF5()
    h = h5min.read(prm.fname_mom_h5)
    assert len(h) == len(mom) * loop.cPrm.nLoop * 16 and set(h) == {"/" + k for k in z.files}
    assert all(h["/" + k].shape == (L[3], 2) and np.array_equal(h["/" + k], z[k]) for k in z.files)


def test_displace_mirror_state_machine(ops, oracle):
    from mugiq_b200.loop import Displace
    from mugiq_b200.params import DISPLACE_TYPE_COVARIANT
    L = (4, 4, 4, 4)
    U = synth.random_gauge(L, seed=39)
    v = synth.random_evecs_np(L, 1, seed=39)[0]
    d = Displace(MugiqLoopParam(gauge=[U[mu] for mu in range(4)]), L)
    vd = dev(v)
    with pytest.raises(MugiqError):
        d.doVectorDisplacement(DISPLACE_TYPE_COVARIANT, vd, 1)
    with pytest.raises(MugiqError, match="Cannot parse"):
        d.setupDisplacement("+w")
    d.setupDisplacement("-z")
    d.doVectorDisplacement(DISPLACE_TYPE_COVARIANT, vd, 1)
    d.doVectorDisplacement(DISPLACE_TYPE_COVARIANT, vd, 2)
    ref = oracle.displace(oracle.displace(v, U, 2, 0, L), U, 2, 0, L)
    assert rel_err(host(vd), ref) < 1e-14
    with pytest.raises(MugiqError, match="Unsupported Displacement"):
        d.doVectorDisplacement(7, vd, 1)
    with pytest.raises(MugiqError, match="Incompatible precision"):
        Displace(MugiqLoopParam(gauge=[U[mu].astype(np.complex64) for mu in range(4)]), L)


def test_full_size_properties(ops):
    """BASELINE.json configs[1] lattice (16^3x32), a few eigenvectors: size-independent properties instead
    of an oracle run — D_- D_+ = 1, linearity in the eigenvector set, minus loop = shifted dagger of plus loop,
    p = 0 projection = spatial sum, sum_x T_1 = sum_n 1/sigma_n."""
    L = (16, 16, 16, 32)
    lat = Lattice(L)
    nEv = 4
    U = synth.random_gauge(L, seed=40)
    gd = ops.gauge_upload(U, L)
    ev = synth.random_evecs_torch(L, nEv, seed=40)
    sig = synth.sigmas(nEv)
    v = ev[0]
    for d in range(4):
        a, b = torch.empty_like(v), torch.empty_like(v)
        ops.displace(a, v, gd, d, 1, L)
        ops.displace(b, a, gd, d, 0, L)
        assert float((b - v).abs().max() / v.abs().max()) < 1e-13
    entries = [(0, 1, 1, 1), (0, 0, 1, 1), (3, 1, 1, 2), (3, 0, 1, 2)]
    nLoop = 7
    full = torch.zeros((nLoop, 16, lat.volume), dtype=torch.complex128, device="cuda")
    ops.loop_accumulate(full, list(ev), sig, gd, entries, L)
    parts = torch.zeros_like(full)
    for n in range(nEv):
        ops.loop_accumulate(parts, [ev[n]], sig[n:n + 1], gd, entries, L, accumulate=n > 0)
    scale = float(full.abs().max())
    assert float((full - parts).abs().max()) / scale < TOL_F64
    assert abs(complex(full[0, 0].sum()) - (1.0 / sig).sum()) < 1e-10 * (1.0 / sig).sum()
    # minus-x loop at x equals +-conj of the plus-x loop at x - mu (Gamma^dag = +-Gamma); same for t, 2 hops
    from oracle import numpy_check as npc
    g = npc.gamma_dense()
    herm = torch.tensor([1.0 if np.allclose(g[G], g[G].conj().T) else -1.0 for G in range(16)], device="cuda")
    for (d, k, ip, im_) in [(0, 1, 1, 2), (3, 1, 3, 5), (3, 2, 4, 6)]:
        idx = np.arange(lat.volume)
        for _ in range(k):
            idx = lat.neighbour_eo(d, 0)[idx]
        idx = torch.from_numpy(idx).cuda()
        want = herm[:, None] * full[ip][:, idx].conj()
        assert float((full[im_] - want).abs().max()) / scale < TOL_F64
    # reorder + p=0 projection = signed spatial sum per time-slice
    mp = torch.empty((lat.V3, 16 * nLoop, L[3]), dtype=torch.complex128, device="cuda")
    ops.reorder_mapgamma(mp, full, 16 * nLoop, nLoop, L)
    ph = ops.phase_matrix([[0, 0, 0]], -1, L)
    dm = ops.momproj(mp, ph, L[3] * 16 * nLoop, 1, lat.V3).reshape(16 * nLoop, L[3])
    ssum = mp.sum(dim=0)
    assert float((dm - ssum).abs().max() / ssum.abs().max()) < TOL_F64


def test_config3_lattice_properties(ops, monkeypatch):
    """BASELINE.json configs[2] lattice and loop set (24^3x48, displacements of length 1..4 in all 8 directions, momenta
    |p|^2 <= 4) with a few eigenvectors.  No oracle at this size: the derived minus loops (identity, one pass over the
    loop buffer) must equal the minus loops computed directly with their own Wilson lines, batches must add up, and
    the one-kernel projection must equal reorder + GEMM."""
    from mugiq_b200.params import parse_disp_entries, which_displace
    L = (24, 24, 24, 48)
    lat = Lattice(L)
    nEv = 3
    U = synth.random_gauge(L, seed=70)
    gd = ops.gauge_upload(U, L)
    ev = synth.random_evecs_torch(L, nEv, seed=70)
    sig = synth.sigmas(nEv)
    _, ds, a, b = parse_disp_entries(synth.UP_TO_4_ENTRIES)
    entries = [which_displace(s) + (x, y) for s, x, y in zip(ds, a, b)]
    plan = ops.LoopPlan(gd, entries, L)
    info = plan.info()
    assert plan.nLoop == 33 and info["computed"] == 17 and info["derived"] == 16
    assert info["wilson_bytes"] == 12 * lat.volume * 144   # plus chains of 2, 3, 4 links in 4 directions
    sym = torch.zeros((33, 16, lat.volume), dtype=torch.complex128, device="cuda")
    plan.accumulate(sym, list(ev[:2]), sig[:2], accumulate=False)
    plan.accumulate(sym, list(ev[2:]), sig[2:], accumulate=True)
    plan.finalize(sym)
    plan.close()
    monkeypatch.setenv("MUGIQ_B200_NO_PM_SYMMETRY", "1")
    direct_plan = ops.LoopPlan(gd, entries, L)
    assert direct_plan.info()["computed"] == 33
    direct = torch.zeros_like(sym)
    direct_plan.accumulate(direct, list(ev), sig, accumulate=False)
    direct_plan.finalize(direct)
    direct_plan.close()
    scale = float(direct.abs().max())
    assert float((sym - direct).abs().max()) / scale < TOL_F64
    assert abs(complex(sym[0, 0].sum()) - (1.0 / sig).sum()) < 1e-10 * (1.0 / sig).sum()
    del direct
    mom = momenta_up_to(4)
    assert len(mom) == 33
    fused = ops.momproj_pos(sym, ops.phase_matrix_eo(mom, -1, L), 33, L)
    mp = torch.empty((lat.V3, 16 * 33, L[3]), dtype=torch.complex128, device="cuda")
    ops.reorder_mapgamma(mp, sym, 16 * 33, 33, L)
    two = ops.momproj(mp, ops.phase_matrix(mom, -1, L), L[3] * 16 * 33, 33, lat.V3).reshape(33, 16 * 33, L[3])
    assert float((fused - two).abs().max() / two.abs().max()) < TOL_F64
