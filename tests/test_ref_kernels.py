"""The reference's OWN CUDA kernels (compiled unmodified against oracle/quda_shim, oracle/_ref/libmugiq_ref.so) on the
B200, against the CPU oracle and against the product's kernels on identical seeded inputs.  This is the closest thing to
"the reference's own QUDA-backed loop kernels" (BASELINE.json north_star) that can run without QUDA: the kernel bodies,
gamma tables, launch geometry and wrappers are the reference's; QUDA's accessors are restated in the shim.
FP64 criterion 1e-12 (norm-relative); pure data movement bit-exact."""
import numpy as np
import pytest
import torch

from conftest import rel_err, TOL_F64, TOL_F32
from mugiq_b200 import synth
from mugiq_b200.params import momenta_up_to

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_kernels
    assert torch.cuda.is_available()
    if not ref_kernels.available():
        pytest.skip("oracle/_ref/libmugiq_ref.so not built (needs /root/reference at build time)")
    return ref_kernels


@pytest.fixture(scope="module")
def ops():
    from mugiq_b200 import ops as o
    return o


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


def cdt(prec):
    return np.complex128 if prec == 8 else np.complex64


@pytest.mark.parametrize("L", [(4, 4, 4, 8), (8, 4, 2, 6), (6, 2, 4, 2)])
@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("prec", [8, 4])
def test_reference_contraction_kernel(ref, ops, oracle, L, order, prec):
    ev = synth.random_evecs_np(L, 3, seed=51).astype(cdt(prec))
    V4 = ev.shape[1]
    tol = TOL_F64 if prec == 8 else TOL_F32
    q = [ref.site_to_quda(dev(ev[i]), order) for i in range(3)]
    assert torch.equal(ref.quda_to_site(q[0], order), dev(ev[0]))
    # the product's layout conversion agrees with the accessor the reference kernels read through
    assert torch.equal(ops.export_spinor(dev(ev[0]), order, L), q[0])
    loop_ref = torch.zeros((16, V4), dtype=q[0].dtype, device="cuda")
    loop_new = torch.zeros_like(loop_ref)
    want = np.zeros((16, V4), dtype=cdt(prec))
    for (a, b, s) in [(0, 1, 0.37), (1, 1, 0.011), (2, 0, 1.9)]:  # accumulating calls, vL != vR and vL == vR
        ref.contract(loop_ref, q[a], q[b], s, L, order)
        ops.contract(loop_new, dev(ev[a]), dev(ev[b]), s, L)
        want = oracle.contract(want, ev[a], ev[b], s, L)
    assert rel_err(host(loop_ref), want) < tol          # oracle vs the reference's kernel
    assert rel_err(host(loop_new), host(loop_ref)) < tol  # product vs the reference's kernel


@pytest.mark.parametrize("L", [(4, 4, 4, 8), (8, 4, 2, 6), (6, 2, 4, 2)])
@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("extended", [True, False])
def test_reference_displacement_kernel(ref, ops, oracle, L, order, extended):
    v = synth.random_evecs_np(L, 1, seed=52)[0]
    U = synth.random_gauge(L, seed=52)
    gd = ops.gauge_upload(U, L)
    vd = dev(v)
    vq = ref.site_to_quda(vd, order)
    for d in range(4):
        for s in (0, 1):
            out_q = torch.full_like(vq, 5.0)
            ref.displace(out_q, vq, gd, d, s, L, order, extended)
            got = host(ref.quda_to_site(out_q, order))
            assert rel_err(got, oracle.displace(v, U, d, s, L)) < 1e-14, (d, s)
            mine = torch.empty_like(vd)
            ops.displace(mine, vd, gd, d, s, L)
            assert rel_err(host(mine), got) < 1e-14, (d, s)


@pytest.mark.parametrize("L", [(4, 4, 4, 8), (8, 4, 2, 6), (6, 2, 4, 2)])
@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("prec", [8, 4])
def test_native_order_kernels_against_reference_kernels(ref, ops, oracle, L, order, prec):
    """mugiq_b200_contract_native / _displace_native work on the very FLOAT2 / FLOAT4 buffers the reference's kernels read
    through FieldOrderCB: same inputs in place, no layout conversion on either side."""
    n = 4
    ev = synth.random_evecs_np(L, 2 * n, seed=54).astype(cdt(prec))
    U = synth.random_gauge(L, seed=54).astype(cdt(prec))
    V4 = ev.shape[1]
    tol = TOL_F64 if prec == 8 else TOL_F32
    q = [ref.site_to_quda(dev(ev[i]), order) for i in range(2 * n)]
    sig = [0.3 + 0.1 * i for i in range(n)]
    loop_ref = torch.zeros((16, V4), dtype=q[0].dtype, device="cuda")
    for i in range(n):
        ref.contract(loop_ref, q[i], q[n + i], sig[i], L, order)
    loop_new = torch.full_like(loop_ref, 2.0)
    ops.contract_native(loop_new, q[:n], q[n:], sig, order, L, accumulate=False)
    assert rel_err(host(loop_new), host(loop_ref)) < tol
    ops.contract_native(loop_new, q[:1], None, sig[:1], order, L, accumulate=True)       # vR = vL, accumulating
    ref.contract(loop_ref, q[0], q[0], sig[0], L, order)
    assert rel_err(host(loop_new), host(loop_ref)) < tol
    gd = dev(U)
    for d in range(4):
        for s in (0, 1):
            outs_ref = [torch.zeros_like(q[0]) for _ in range(3)]
            for i in range(3):
                ref.displace(outs_ref[i], q[i], gd, d, s, L, order, True)
            outs = [torch.full_like(q[0], 9.0) for _ in range(3)]
            ops.displace_native(outs, q[:3], gd, d, s, order, L)
            for i in range(3):
                assert rel_err(host(outs[i]), host(outs_ref[i])) < (1e-14 if prec == 8 else 1e-6), (d, s, i)
            assert rel_err(host(ref.quda_to_site(outs[0], order)), oracle.displace(ev[0], U, d, s, L)) < (1e-14 if prec == 8 else 1e-6)


@pytest.mark.parametrize("L", [(4, 4, 4, 8), (8, 2, 6, 4), (6, 4, 2, 3)])
@pytest.mark.parametrize("prec", [8, 4])
def test_reference_reorder_kernel(ref, ops, oracle, L, prec):
    nLoop = 3
    V4 = int(np.prod(L))
    rng = np.random.default_rng(5)
    pos = (rng.standard_normal((nLoop, 16, V4)) + 1j * rng.standard_normal((nLoop, 16, V4))).astype(cdt(prec))
    out_ref = torch.zeros(V4 * 16 * nLoop, dtype=dev(pos).dtype, device="cuda")
    ref.reorder_mapgamma(out_ref, dev(pos), nLoop, L)
    out_new = torch.zeros_like(out_ref)
    ops.reorder_mapgamma(out_new, dev(pos), 16 * nLoop, nLoop, L)
    assert torch.equal(out_new, out_ref)  # data movement and a sign: bit-exact
    assert np.array_equal(host(out_ref), oracle.reorder_mapgamma(pos, nLoop, L).reshape(-1))


@pytest.mark.parametrize("ftsign", [-1, 1])
def test_reference_phase_kernel(ref, ops, oracle, ftsign):
    L = (4, 6, 8, 4)
    mom = momenta_up_to(4)
    ph_ref = host(ref.phase_matrix(mom, ftsign, L))
    ph_new = host(ops.phase_matrix(mom, ftsign, L))
    # cos(2.0*PI*phi) with PI = 2*asin(1) (include/util_mugiq.h:7) against sincospi on the same argument
    assert np.abs(ph_new - ph_ref).max() < 5e-15
    assert np.abs(oracle.phase_matrix(mom, ftsign, L) - ph_ref).max() < 5e-15


@pytest.mark.parametrize("L,entries", [((4, 4, 4, 8), synth.ONE_HOP_ENTRIES + ";+z:2,3;-t:1,2"), ((8, 2, 4, 4), "+x:1,3;-x:1,3;-y:2")])
def test_reference_loop_nest(ref, ops, oracle, L, entries):
    """Whole position-space loop buffer: the reference's kernels in the reference's loop order vs the oracle and vs the
    product's fused path, then the projection chain reorder -> phase -> GEMM vs the fused projection."""
    from mugiq_b200.loop import Loop_Mugiq, Eigsolve
    from mugiq_b200.params import MugiqLoopParam
    nEv = 5
    ev = synth.random_evecs_np(L, nEv, seed=53)
    sig = synth.sigmas(nEv)
    U = synth.random_gauge(L, seed=53)
    prm = MugiqLoopParam(gauge=[U[mu] for mu in range(4)])
    prm.set_displacements(entries)
    mom = momenta_up_to(2)
    prm.set_momenta(mom)
    loop = Loop_Mugiq(prm, Eigsolve([dev(ev[i]) for i in range(nEv)], sig, L))
    loop.computeCoarseLoop()
    ents = loop.cPrm.entries()
    gd = ops.gauge_upload(U, L)
    pos_ref = ref.compute_loop([ref.site_to_quda(dev(ev[i]), 2) for i in range(nEv)], sig, gd, ents, L)
    assert rel_err(host(pos_ref), oracle.compute_loop(ev, sig, U, ents, L)) < TOL_F64
    assert rel_err(loop.dataPos.numpy(), host(pos_ref)) < TOL_F64
    # momentum projection of the reference: convertIdxOrder_mapGamma + phase matrix + ZGEMM (torch.matmul = cuBLAS)
    nLoop, V3, Lt = pos_ref.shape[0], L[0] * L[1] * L[2], L[3]
    mp = torch.zeros(V3 * 16 * nLoop * Lt, dtype=torch.complex128, device="cuda")
    ref.reorder_mapgamma(mp, pos_ref, nLoop, L)
    ph = ref.phase_matrix(mom, -1, L)                              # [Nmom, V3]
    dm = torch.matmul(ph, mp.reshape(V3, 16 * nLoop * Lt))         # dataMom[im][idata][t]
    assert rel_err(loop.dataMom.numpy().reshape(len(mom), -1), host(dm)) < TOL_F64


def test_reference_kernel_only_timing(ref, ops):
    """bench.py's `reference_gpu.kernel_only`: the reference's loop nest with CUDA events around its kernels only (argument
    structs pre-staged).  Checks the bookkeeping: one contraction per (eigenvector, loop), `stop` hops per eigenvector and
    entry, positive times."""
    L, nEv = (4, 4, 4, 8), 3
    ev = synth.random_evecs_np(L, nEv, seed=54)
    U = synth.random_gauge(L, seed=54)
    entries = [(0, 1, 1, 1), (3, 0, 2, 3)]
    gd = ops.gauge_upload(U, L)
    k = ref.kernel_only_ms([ref.site_to_quda(dev(ev[i]), 2) for i in range(nEv)], synth.sigmas(nEv), gd, entries, L)
    assert k["launches"] == nEv * (1 + 1 + 2) + nEv * (1 + 3)   # contractions + hops
    assert k["contract_ms"] > 0 and k["displace_ms"] > 0 and k["field_copies_ms"] > 0 and abs(k["ms"] - k["contract_ms"] - k["displace_ms"]) < 1e-6
