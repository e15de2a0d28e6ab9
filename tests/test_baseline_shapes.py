"""Parity of the CUDA path against the CPU oracle AND the reference's own kernels (oracle/_ref) on the lattices
BASELINE.json names, through the public Loop_Mugiq interface with its default batching, with the minus-from-plus
derivation on and off: 16^3x32 (configs[1]), 24^3x48 with displacements 1..4 and |p|^2 <= 4 (configs[2]), 32^3x64
ultra-local (configs[3]) and the extended 48^3 time slab of configs[4] (interior-only compute, set_t_range).
A few eigenvectors each: the oracle does ~3e7 contractions/s, so every case takes seconds.
FP64 criterion: max|delta| / max|ref| <= 1e-12 per buffer (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from conftest import rel_err, TOL_F64
from mugiq_b200 import synth
from mugiq_b200.lattice import Lattice
from mugiq_b200.params import MugiqLoopParam, momenta_up_to, parse_disp_entries, which_displace

pytestmark = pytest.mark.gpu


def entry_list(text):
    if not text:
        return []
    _, ds, a, b = parse_disp_entries(text)
    return [which_displace(s) + (x, y) for s, x, y in zip(ds, a, b)]


def rel_err_t(a, b):
    """max|a-b| / max|b| of two device tensors (5.6 GB buffers stay on the GPU)."""
    return float((a - b).abs().max() / b.abs().max())


def reference_kernels_loop(ev, sig, U, entries, L):
    """dataPos from the reference's own kernels and wrappers (oracle/_ref), or None where the library was not built."""
    from oracle import ref_kernels as ref
    if not ref.available():
        return None
    evq = [ref.site_to_quda(torch.from_numpy(ev[n]).cuda(), 2) for n in range(ev.shape[0])]
    gd = torch.from_numpy(U).cuda() if entries else None
    out = ref.compute_loop(evq, sig, gd, entries, L, order=2)
    del evq, gd
    return out


def run_loop(ev, sig, U, entries_text, mom, L, **kw):
    from mugiq_b200.loop import Loop_Mugiq, Eigsolve
    prm = MugiqLoopParam(gauge=[U[mu] for mu in range(4)] if U is not None else None)
    if entries_text:
        prm.set_displacements(entries_text)
    if mom is not None:
        prm.set_momenta(mom)
    loop = Loop_Mugiq(prm, Eigsolve([torch.from_numpy(ev[n]).cuda() for n in range(ev.shape[0])], sig, L),
                      copy_pos_to_host=False, **kw)
    loop.computeCoarseLoop()
    return loop


@pytest.mark.parametrize("symmetry", [True, False])
def test_config2_16x16x16x32(oracle, monkeypatch, symmetry):
    """configs[1]: ultra-local + the 8 one-hop loops, 7 momenta, 8 eigenvectors, default evec_batch."""
    if not symmetry:
        monkeypatch.setenv("MUGIQ_B200_NO_PM_SYMMETRY", "1")
    from oracle import numpy_check as npc
    L, nEv = (16, 16, 16, 32), 8
    ev = synth.random_evecs_np(L, nEv, seed=201)
    sig = synth.sigmas(nEv)
    U = synth.random_gauge(L, seed=201)
    mom = momenta_up_to(1)
    loop = run_loop(ev, sig, U, synth.ONE_HOP_ENTRIES, mom, L)
    assert loop._plan.info()["computed"] == (5 if symmetry else 9)
    ref = oracle.compute_loop(ev, sig, U, loop.cPrm.entries(), L)
    assert rel_err(loop.dataPos_d.cpu().numpy(), ref) < TOL_F64
    ref_mom = npc.momentum_projection_mm(ref, mom, -1, L)
    assert rel_err(loop.dataMom.numpy(), ref_mom) < TOL_F64
    rk = reference_kernels_loop(ev, sig, U, loop.cPrm.entries(), L)
    if rk is not None:
        assert rel_err_t(loop.dataPos_d, rk) < TOL_F64


def test_config3_24x24x24x48(oracle):
    """configs[2]: displacements of length 1..4 in all 8 directions (33 loops, 17 computed in 5 launch groups of the
    fused kernel on runs that are 2 2/3 lattice rows long), 33 momenta, 2 eigenvectors."""
    from oracle import numpy_check as npc
    L, nEv = (24, 24, 24, 48), 2
    ev = synth.random_evecs_np(L, nEv, seed=202)
    sig = synth.sigmas(nEv)
    U = synth.random_gauge(L, seed=202)
    mom = momenta_up_to(4)
    loop = run_loop(ev, sig, U, synth.UP_TO_4_ENTRIES, mom, L)
    assert loop.cPrm.nLoop == 33 and len(mom) == 33
    ref = oracle.compute_loop(ev, sig, U, loop.cPrm.entries(), L)
    ref_d = torch.from_numpy(ref).cuda()
    assert rel_err_t(loop.dataPos_d, ref_d) < TOL_F64
    del ref_d
    # projection: ultra-local, a plus and a derived minus loop of length 1, the longest t loops
    slots = [0, 1, 5, 28, 32]
    ref_mom = npc.momentum_projection_mm(ref, mom, -1, L, loops=slots)
    got = loop.dataMom.numpy().reshape(33, 33, 16, L[3])[:, slots].reshape(33, 16 * len(slots), L[3])
    assert rel_err(got, ref_mom) < TOL_F64
    del ref
    rk = reference_kernels_loop(ev, sig, U, loop.cPrm.entries(), L)
    if rk is not None:
        assert rel_err_t(loop.dataPos_d, rk) < TOL_F64


def test_config4_32x32x32x64_ultralocal(oracle):
    """configs[3]: the ultra-local 16-gamma loop on 32^3x64 (runs of 128 sites per parity), 2 eigenvectors, p = 0."""
    from oracle import numpy_check as npc
    L, nEv = (32, 32, 32, 64), 2
    ev = synth.random_evecs_np(L, nEv, seed=203)
    sig = synth.sigmas(nEv)
    loop = run_loop(ev, sig, None, "", momenta_up_to(0), L)
    ref = oracle.compute_loop(ev, sig, None, [], L)
    assert rel_err(loop.dataPos_d.cpu().numpy(), ref) < TOL_F64
    assert rel_err(loop.dataMom.numpy(), npc.momentum_projection_mm(ref, momenta_up_to(0), -1, L)) < TOL_F64
    rk = reference_kernels_loop(ev, sig, None, [], L)
    if rk is not None:
        assert rel_err_t(loop.dataPos_d, rk) < TOL_F64


@pytest.mark.parametrize("symmetry", [True, False])
def test_config5_time_slab_48x48x48(oracle, monkeypatch, symmetry):
    """configs[4]: one rank's time slab of the 48^3x96 lattice on 8 GPUs, 12 slices extended by H = 2 halo slices on each
    side (48^3x16), plan restricted to the interior with set_t_range: the kernels read the halos and compute the
    interior only.  The oracle runs on the extended slab as a periodic lattice; interior values agree because no loop
    reaches further than one slice.  With the minus-from-plus derivation the -t loop of the lowest interior slice comes
    from the loop-buffer halo the T split fetches from the neighbour (tests/test_tsplit.py), so it is excluded here."""
    if not symmetry:
        monkeypatch.setenv("MUGIQ_B200_NO_PM_SYMMETRY", "1")
    from mugiq_b200 import ops
    L, H, Tl, nEv = (48, 48, 48, 16), 2, 12, 2
    lat = Lattice(L)
    ev = synth.random_evecs_np(L, nEv, seed=204)
    sig = synth.sigmas(nEv)
    U = synth.random_gauge(L, seed=204)
    entries = entry_list(synth.ONE_HOP_ENTRIES)
    gd = ops.gauge_upload(U, L)
    plan = ops.LoopPlan(gd, entries, L)
    plan.set_t_range(H, H + Tl)
    pos = torch.full((9, 16, lat.volume), float("nan"), dtype=torch.complex128, device="cuda")
    evd = [torch.from_numpy(ev[n]).cuda() for n in range(nEv)]
    plan.accumulate(pos, evd[:1], sig[:1], accumulate=False)
    # accumulate mode reads the previous partial sums: interior only, the halo slices still hold the NaN fill
    plan.accumulate(pos, evd[1:], sig[1:], accumulate=True)
    t_of = torch.from_numpy(lat.coords_eo()[:, 3]).cuda()
    interior = (t_of >= H) & (t_of < H + Tl)
    computed = [0, 1, 3, 5, 7] if symmetry else list(range(9))
    assert not torch.isnan(pos[computed][:, :, interior].real).any()
    assert torch.isnan(pos[computed][:, :, ~interior].real).all()   # nothing outside the range was written
    ref = torch.from_numpy(oracle.compute_loop(ev, sig, U, entries, L)).cuda()
    for iL in computed:
        assert rel_err_t(pos[iL][:, interior], ref[iL][:, interior]) < TOL_F64
    if symmetry:
        pos[:, :, ~interior] = ref[:, :, ~interior]   # what the neighbours' loop-buffer halo would deliver
        plan.finalize(pos)
        for iL in range(9):
            assert rel_err_t(pos[iL][:, interior], ref[iL][:, interior]) < TOL_F64
    rk = reference_kernels_loop(ev, sig, U, entries, L)
    if rk is not None:
        for iL in computed:
            assert rel_err_t(pos[iL][:, interior], rk[iL][:, interior]) < TOL_F64
    plan.close()


def test_resident_batches_follow_changed_sigma_and_refilled_buffers(oracle):
    """Loop_Mugiq keeps the argument tables of a device-resident batch between calls.  Running again after the caller
    changed eVals_sigma, refilled the same device buffers with new eigenvectors, or swapped a tensor must give the new
    result (the reference re-reads sigma and the fields on every call, lib/loop_mugiq.cpp:478-483)."""
    from mugiq_b200.loop import Loop_Mugiq, Eigsolve
    L, nEv = (8, 4, 4, 8), 6
    ev = synth.random_evecs_np(L, 2 * nEv, seed=205)
    sig = synth.sigmas(nEv)
    U = synth.random_gauge(L, seed=205)
    prm = MugiqLoopParam(gauge=[U[mu] for mu in range(4)])
    prm.set_displacements("+x:1;-t:1,2")
    es = Eigsolve([torch.from_numpy(ev[n]).cuda() for n in range(nEv)], sig, L)
    loop = Loop_Mugiq(prm, es, evec_batch=4, copy_pos_to_host=False)
    entries = loop.cPrm.entries()
    loop.computeCoarseLoop()
    assert rel_err(loop.dataPos_d.cpu().numpy(), oracle.compute_loop(ev[:nEv], sig, U, entries, L)) < TOL_F64
    # 1. new sigma values, same buffers
    sig2 = sig * 3.0 + 0.5
    es.eVals_sigma = [float(x) for x in sig2]
    loop.computeCoarseLoop()
    assert rel_err(loop.dataPos_d.cpu().numpy(), oracle.compute_loop(ev[:nEv], sig2, U, entries, L)) < TOL_F64
    # 2. same buffers refilled with other eigenvectors
    for n in range(nEv):
        es.eVecs[n].copy_(torch.from_numpy(ev[nEv + n]))
    loop.computeCoarseLoop()
    assert rel_err(loop.dataPos_d.cpu().numpy(), oracle.compute_loop(ev[nEv:], sig2, U, entries, L)) < TOL_F64
    # 3. an interior tensor of a batch replaced by another allocation
    es.eVecs[2] = torch.from_numpy(ev[0]).cuda()
    mixed = np.concatenate([ev[nEv:nEv + 2], ev[0:1], ev[nEv + 3:]])
    loop.computeCoarseLoop()
    assert rel_err(loop.dataPos_d.cpu().numpy(), oracle.compute_loop(mixed, sig2, U, entries, L)) < TOL_F64


@pytest.mark.parametrize("L,entries_text,p2", [((16, 16, 16, 32), synth.ONE_HOP_ENTRIES, 1), ((24, 24, 24, 48), synth.UP_TO_4_ENTRIES, 0),
                                              ((8, 4, 4, 8), "+x:1,3;-y:2;+t:1;-t:1;+z:2", 1), ((4, 4, 4, 8), "", 1)])
def test_quda_float2_eigenvectors_are_staged_directly(oracle, L, entries_text, p2):
    """Eigenvectors handed over in QUDA's native FLOAT2 order ([parity][spin*3+colour][x_cb], what FieldOrderCB gives the
    reference's kernels, lib/mugiq_contract_kernels.cu:82-83): the fused kernel fetches them as TMA tensor boxes, no
    conversion pass.  Same oracle, same tolerance; device-resident and host-streamed (feed) eigenvectors."""
    from mugiq_b200.loop import Loop_Mugiq, Eigsolve
    from oracle import ref_kernels as rk
    from oracle import numpy_check as npc
    nEv = 5 if L[0] <= 8 else 3
    ev = synth.random_evecs_np(L, nEv, seed=206)
    sig = synth.sigmas(nEv)
    U = synth.random_gauge(L, seed=206) if entries_text else None
    mom = momenta_up_to(p2)
    prm = MugiqLoopParam(gauge=[U[mu] for mu in range(4)] if U is not None else None)
    if entries_text:
        prm.set_displacements(entries_text)
    prm.set_momenta(mom)
    evq = [rk.site_to_quda(torch.from_numpy(ev[n]).cuda(), 2) for n in range(nEv)]
    loop = Loop_Mugiq(prm, Eigsolve(evq, sig, L, field_order=2), evec_batch=2, copy_pos_to_host=False)
    loop.computeCoarseLoop()
    ref = oracle.compute_loop(ev, sig, U, entry_list(entries_text), L)
    ref_d = torch.from_numpy(ref).cuda()
    assert rel_err_t(loop.dataPos_d, ref_d) < TOL_F64
    if L[0] <= 16:
        assert rel_err(loop.dataMom.numpy(), npc.momentum_projection_mm(ref, mom, -1, L)) < TOL_F64
    # the same fields from pinned host memory through the streamed feed (FLOAT2 staging batches, no conversion)
    evq_h = [v.cpu().pin_memory() for v in evq]
    loop_h = Loop_Mugiq(prm, Eigsolve(evq_h, sig, L, field_order=2), stream_batch=2, copy_pos_to_host=False)
    loop_h.computeCoarseLoop()
    assert rel_err_t(loop_h.dataPos_d, ref_d) < TOL_F64


def test_quda_float2_single_precision(oracle):
    from mugiq_b200.loop import Loop_Mugiq, Eigsolve
    from oracle import ref_kernels as rk
    L, nEv = (8, 4, 4, 8), 4
    ev = synth.random_evecs_np(L, nEv, seed=207)
    sig = synth.sigmas(nEv)
    U = synth.random_gauge(L, seed=207)
    prm = MugiqLoopParam(gauge=[U[mu].astype(np.complex64) for mu in range(4)])
    prm.set_displacements("+x:1;-z:1,2;+t:2")
    evq = [rk.site_to_quda(torch.from_numpy(ev[n].astype(np.complex64)).cuda(), 2) for n in range(nEv)]
    loop = Loop_Mugiq(prm, Eigsolve(evq, sig, L, field_order=2), copy_pos_to_host=False)
    loop.computeCoarseLoop()
    ref = oracle.compute_loop(ev, sig, U, entry_list("+x:1;-z:1,2;+t:2"), L)
    assert rel_err(loop.dataPos_d.cpu().numpy(), ref) < 2e-5
