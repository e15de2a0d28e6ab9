"""Lattice-T split (mugiq_b200/tsplit.py): geometry and halo exchange on the CPU (gloo, world_size 2), and on the GPU
the split computation against the oracle's result on the global lattice."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import rel_err, TOL_F64
from mugiq_b200 import synth
from mugiq_b200.lattice import Lattice
from mugiq_b200.tsplit import TSplit

L = (4, 2, 6, 8)


def _global_field(seed=3, ncomp=12, nb=3):
    rng = np.random.default_rng(seed)
    V4 = int(np.prod(L))
    return rng.standard_normal((nb, V4, ncomp)) + 1j * rng.standard_normal((nb, V4, ncomp))


def _interior_of_global(ts, field, site_dim):
    """this rank's owned time-slices cut out of a global even/odd array"""
    return ts.interior(ts.global_slab(field, site_dim=site_dim), site_dim=site_dim)


def test_geometry_and_parity_bookkeeping():
    lat = Lattice(L)
    coords = lat.coords_eo()  # [V4, 4] in global even/odd order
    tag = (coords[:, 0] + 10 * coords[:, 1] + 100 * coords[:, 2] + 1000 * coords[:, 3]).astype(np.float64)
    for world in (1, 2, 4):
        for rank in range(world):
            ts = TSplit(L, rank, world, max_t_disp=1)
            assert ts.H == 2 and ts.L_ext == (4, 2, 6, L[3] // world + 4)
            slab = ts.global_slab(tag, site_dim=0)
            ext = Lattice(ts.L_ext).coords_eo()
            t_glob = (ext[:, 3] - ts.H + ts.t0) % L[3]
            want = ext[:, 0] + 10 * ext[:, 1] + 100 * ext[:, 2] + 1000 * t_glob
            assert np.array_equal(slab, want)  # a site of the extended slab IS the global site with shifted t
            inner = ts.interior(torch.from_numpy(slab), site_dim=0).numpy()
            loc = Lattice(ts.L_loc).coords_eo()
            assert np.array_equal(inner, loc[:, 0] + 10 * loc[:, 1] + 100 * loc[:, 2] + 1000 * (loc[:, 3] + ts.t0))
    with pytest.raises(ValueError):
        TSplit(L, 0, 3, 1)
    with pytest.raises(ValueError):
        TSplit((4, 4, 4, 6), 0, 2, 1)  # local T = 3 is odd
    with pytest.raises(ValueError):
        TSplit(L, 0, 4, 3)             # halo 4 > local T = 2


def test_single_rank_extension_is_the_periodic_wrap():
    g = torch.from_numpy(_global_field())
    ts = TSplit(L, 0, 1, max_t_disp=2)
    want = ts.global_slab(g, site_dim=1)
    assert torch.equal(ts.extend(g), want)
    # in-place form: the caller's fields already have the halo slices allocated, only the halos are (re)written
    stored = want.clone()
    v = stored.reshape(-1, 2, ts.Tl + 2 * ts.H, ts.V3h, 12)
    v[:, :, :ts.H] = 0
    v[:, :, ts.H + ts.Tl:] = 0
    out = ts.finish_extend(ts.begin_extend(list(stored)))
    assert all(o.data_ptr() == s.data_ptr() for o, s in zip(out, stored)) and torch.equal(torch.stack(out), want)


def test_slab_gauge_generator_is_one_global_su3_field():
    """Slice-keyed generator: unitary, det = 1, and two slabs agree wherever they hold the same global time-slice."""
    Lg = (4, 4, 2, 8)
    a = synth.random_gauge_slab_torch(Lg, [6, 7, 0, 1, 2, 3], seed=3, device="cpu")   # rank 0 of 4 ranks... with halo 2
    b = synth.random_gauge_slab_torch(Lg, [0, 1, 2, 3, 4, 5], seed=3, device="cpu")
    V3h = 4 * 4 * 2 // 2
    for mu in range(4):
        A = a[mu].reshape(2, 6, V3h, 3, 3)
        B = b[mu].reshape(2, 6, V3h, 3, 3)
        assert torch.equal(A[:, 2:6], B[:, 0:4])
        eye = torch.eye(3, dtype=torch.complex128)
        assert (A @ A.conj().transpose(-1, -2) - eye).abs().max() < 1e-13
        assert (torch.linalg.det(A) - 1).abs().max() < 1e-13


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.from_numpy(_global_field())
    ts = TSplit(L, rank, world, max_t_disp=2)
    inner = _interior_of_global(ts, g, 1).contiguous()
    ext = ts.extend(inner, group=dist.group.WORLD)          # halos arrive from the neighbours
    ok = torch.equal(ext, ts.global_slab(g, site_dim=1))
    mom = torch.full((2, 3, ts.Tl), float(rank), dtype=torch.complex128) + torch.arange(ts.Tl)
    gathered = ts.gather_time(mom, group=dist.group.WORLD)
    want = torch.cat([torch.full((2, 3, ts.Tl), float(r), dtype=torch.complex128) + torch.arange(ts.Tl) for r in range(world)], -1)
    ok = ok and torch.equal(gathered, want)
    # one-sided extension: only the requested halo is filled
    up_only = ts.finish_extend(ts.begin_extend(inner, group=dist.group.WORLD, lower=False, upper=True))
    full = ts.global_slab(g, site_dim=1).reshape(-1, 2, ts.Tl + 2 * ts.H, ts.V3h, 12)
    got = up_only.reshape(full.shape)
    ok = ok and torch.equal(got[:, :, ts.H:], full[:, :, ts.H:])
    # exactly the slices that are read: one above, one below (H = 2 is only the parity-preserving offset)
    part = ts.finish_extend(ts.begin_extend(inner, group=dist.group.WORLD, lower=1, upper=1)).reshape(full.shape)
    ok = ok and torch.equal(part[:, :, ts.H - 1:ts.H + ts.Tl + 1], full[:, :, ts.H - 1:ts.H + ts.Tl + 1])
    # loop-buffer halo: the lower halo slices of the chosen slots become the top interior slices of the rank below
    rng = np.random.default_rng(7)
    V4g = int(np.prod(L))
    pos_g = torch.from_numpy(rng.standard_normal((4, 16, V4g)) + 1j * rng.standard_normal((4, 16, V4g)))
    pos_ext = ts.global_slab(pos_g, site_dim=2).clone()
    pv = pos_ext.reshape(4, 16, 2, ts.Tl + 2 * ts.H, ts.V3h)
    pv[:, :, :, :ts.H] = 0
    ts.exchange_loop_halo(pos_ext, [1, 3], group=dist.group.WORLD, depth=1)
    want_ext = ts.global_slab(pos_g, site_dim=2).reshape(pv.shape)
    ok = ok and torch.equal(pv[[1, 3]][:, :, :, ts.H - 1:ts.H], want_ext[[1, 3]][:, :, :, ts.H - 1:ts.H]) and bool((pv[[1, 3]][:, :, :, :ts.H - 1] == 0).all())
    ts.exchange_loop_halo(pos_ext, [1, 3], group=dist.group.WORLD)
    ok = ok and torch.equal(pv[[1, 3]][:, :, :, :ts.H], want_ext[[1, 3]][:, :, :, :ts.H]) and bool((pv[[0, 2]][:, :, :, :ts.H] == 0).all())
    # the loop file written by all time ranks at once (lib/loop_mugiq.cpp:561-572,624): rank 0 lays it out, everybody
    # writes its own rows; compared with the serial file of the gathered data
    from mugiq_b200 import h5lite, h5min

    class FakeLoop:  # what h5lite reads from a Loop_Mugiq: {(momentum, displacement tag, gamma name): [locT] complex}
        def __init__(self, t0, nt):
            self.t0, self.nt = t0, nt

        def momentum_loops(self):
            return {((px, 0, 1), tag, gname): (np.arange(self.t0, self.t0 + self.nt) + 10 * px) * (1 + 2j * ig)
                    for px in (-1, 0) for tag in ("disp_0", "disp_+t_10") for ig, gname in enumerate(("G0", "G15"))}

    path = os.path.join(out_dir, "loops_parallel.h5")
    h5lite.write_momentum_loops_time_ranks(path, FakeLoop(ts.t0, ts.Tl), rank, world, dist.barrier)
    if rank == 0:
        h5lite.write_momentum_loops(os.path.join(out_dir, "loops_serial.h5"), FakeLoop(0, L[3]))
        ok = ok and open(path, "rb").read() == open(os.path.join(out_dir, "loops_serial.h5"), "rb").read()
        ok = ok and len(h5min.read(path)) == 8
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(ok)]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_halo_exchange_gloo(tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(int(np.load(tmp_path / f"ok{r}.npy")[0]) == 1 for r in range(world))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("sides", [(True, True), (False, True), (True, False), (1, 1)])
def test_peer_halo_push_single_rank(mode, sides):
    """NVLink peer mode with one rank: the 'neighbours' are the rank's own allocation, so the pushed halos must be the
    periodic wrap - by the copy engines (mode 0) and by the SM push kernel (mode 1), for a batch in the middle of the
    slab array, and only the requested sides are written."""
    from mugiq_b200 import ops
    lower, upper = sides
    g = torch.from_numpy(_global_field(nb=5)).cuda()
    ts = TSplit(L, 0, 1, max_t_disp=2)
    want = ts.global_slab(g, site_dim=1)                       # [5, V4_ext, 12]
    buf = ops.PeerBuffer(want.numel() * 16)
    try:
        slabs = buf.tensor(tuple(want.shape), torch.complex128)
        slabs.copy_(want)
        v = slabs.reshape(5, 2, ts.Tl + 2 * ts.H, ts.V3h, 12)
        v[:, :, :ts.H] = 7.0
        v[:, :, ts.H + ts.Tl:] = 7.0
        ts.attach_peers(slabs, buf.ptr, buf.ptr, mode=mode)
        out = ts.finish_extend(ts.begin_extend(list(slabs[1:4]), lower=lower, upper=upper))
        torch.cuda.synchronize()
        assert len(out) == 3 and out[0].data_ptr() == slabs[1].data_ptr()
        w = want.reshape(v.shape)
        H, Tl = ts.H, ts.Tl
        lo = H if lower is True else int(lower)
        up = H if upper is True else int(upper)
        assert torch.equal(v[1:4, :, H - lo:H + Tl + up], w[1:4, :, H - lo:H + Tl + up])       # interior + the requested slices
        assert bool((v[1:4, :, :H - lo] == 7.0).all()) and bool((v[1:4, :, H + Tl + up:] == 7.0).all())  # nothing else written
        assert bool((v[0, :, :H] == 7.0).all()) and bool((v[4, :, H + Tl:] == 7.0).all())     # other vectors untouched
    finally:
        del slabs, v
        buf.free()


@pytest.mark.gpu
def test_tsplit_with_device_slab_links(oracle, monkeypatch):
    """A rank may hand Loop_Mugiq its extended slab of links already on the device (bench.py does, for lattices whose
    global field is too large to replicate on the host): same loops as from the replicated host field."""
    from mugiq_b200.loop import Loop_Mugiq, Eigsolve
    from mugiq_b200.params import MugiqLoopParam
    Lg, world, nEv = (4, 4, 2, 8), 2, 3
    Ug = synth.random_gauge_slab_torch(Lg, list(range(Lg[3])), seed=9, device="cuda")
    U = np.stack([u.cpu().numpy() for u in Ug])
    ev = synth.random_evecs_np(Lg, nEv, seed=62)
    sig = synth.sigmas(nEv)
    entries = [(3, 1, 1, 2), (2, 0, 1, 1)]   # no derived minus-t loop: no loop-buffer halo in this test
    ref = oracle.compute_loop(ev, sig, U, entries, Lg)
    evg = torch.from_numpy(ev).cuda()
    parts = []
    for rank in range(world):
        ts = TSplit(Lg, rank, world, max_t_disp=2)
        slab = synth.random_gauge_slab_torch(Lg, [(ts.t0 - ts.H + i) % Lg[3] for i in range(ts.Tl + 2 * ts.H)], seed=9, device="cuda")
        monkeypatch.setattr(ts, "begin_extend", lambda vecs, group=None, device=None, lower=True, upper=True: len(vecs))
        monkeypatch.setattr(ts, "finish_extend", lambda n, ts=ts: ts.global_slab(evg[:n], site_dim=1))
        prm = MugiqLoopParam(gauge=slab)
        prm.set_displacements("+t:1,2;-z:1")
        prm.doMomProj = False
        inner = _interior_of_global(ts, evg, 1).contiguous()
        loop = Loop_Mugiq(prm, Eigsolve(list(inner), sig, ts.L_loc), tsplit=ts, stream_batch=nEv, group=object())
        loop.computeCoarseLoop()
        parts.append(loop.dataPos.numpy().reshape(ref.shape[0], 16, 2, ts.Tl, ts.V3h))
    assert rel_err(np.concatenate(parts, axis=3).reshape(ref.shape), ref) < TOL_F64


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 4])
@pytest.mark.parametrize("symmetric", [True, False])
def test_tsplit_loop_matches_global_oracle(oracle, world, symmetric, monkeypatch):
    """Each (virtual) rank computes the INTERIOR of its time-slab with the fused kernels on the extended lattice; the
    stitched position-space buffer and the gathered momentum-space buffer equal the oracle's on the global lattice.
    The eigenvector halos are cut from the global field here (one process) and the loop-buffer halo of the derived
    minus-t loops is passed between the virtual ranks by hand; the NCCL/gloo exchanges themselves are covered above.
    symmetric=False (MUGIQ_B200_NO_PM_SYMMETRY=1): every loop is computed directly, both eigenvector halos are read."""
    if not symmetric:
        monkeypatch.setenv("MUGIQ_B200_NO_PM_SYMMETRY", "1")
    from mugiq_b200.loop import Loop_Mugiq, Eigsolve
    from mugiq_b200.params import MugiqLoopParam, momenta_up_to
    from oracle import numpy_check as npc
    Lg = (4, 4, 2, 8)
    nEv = 5
    ev = synth.random_evecs_np(Lg, nEv, seed=61)
    sig = synth.sigmas(nEv)
    U = synth.random_gauge(Lg, seed=61)
    entries_str = "+t:1,2;-t:1,2;+x:1;-y:2;-t:1"
    entries = [(3, 1, 1, 2), (3, 0, 1, 2), (0, 1, 1, 1), (1, 0, 2, 2), (3, 0, 1, 1)]
    mom = momenta_up_to(1)
    ref = oracle.compute_loop(ev, sig, U, entries, Lg)
    ref_mom = npc.momentum_projection(ref, mom, -1, Lg)
    evg = torch.from_numpy(ev).cuda()
    H = 2
    sent = {}  # rank -> the loop-buffer slices it sends upwards (recorded in pass 0, delivered in pass 1)

    def run(rank, deliver, copy_pos_to_host=True):
        ts = TSplit(Lg, rank, world, max_t_disp=2)

        def begin(vecs, group=None, device=None, lower=True, upper=True):
            return len(vecs), lower, upper

        def finish(h):
            n, lower, upper = h
            ext = ts.global_slab(evg[:n], site_dim=1).clone()
            v = ext.reshape(n, 2, ts.Tl + 2 * H, ts.V3h, 12)
            # halo slices the plan said it does not read must not influence the result
            lo = H if lower is True else int(lower)
            up = H if upper is True else int(upper)
            v[:, :, :H - lo] = float("nan")
            v[:, :, H + ts.Tl + up:] = float("nan")
            return ext

        def loop_halo(dataPosExt, slots, group=None, depth=None):
            d = H if depth is None else int(depth)
            v = dataPosExt.reshape(dataPosExt.shape[0], 16, 2, ts.Tl + 2 * H, ts.V3h)
            idx = torch.as_tensor(list(slots), device=dataPosExt.device)
            sent[rank] = v[idx][:, :, :, H + ts.Tl - d:H + ts.Tl].clone()
            if deliver:
                v[idx, :, :, H - d:H] = sent[(rank - 1) % world]

        monkeypatch.setattr(ts, "begin_extend", begin)
        monkeypatch.setattr(ts, "finish_extend", finish)
        monkeypatch.setattr(ts, "exchange_loop_halo", loop_halo)
        monkeypatch.setattr(ts, "gather_time", lambda m, group=None: m)
        prm = MugiqLoopParam(gauge=[U[mu] for mu in range(4)])
        prm.set_displacements(entries_str)
        prm.set_momenta(mom)
        inner = _interior_of_global(ts, evg, 1).contiguous()
        loop = Loop_Mugiq(prm, Eigsolve(list(inner), sig, ts.L_loc), tsplit=ts, stream_batch=nEv,
                          group=object() if world > 1 else None, copy_pos_to_host=copy_pos_to_host)
        loop.computeCoarseLoop()
        return ts, loop

    # pass 0 records what every rank would send (its plus-t loops do not depend on the loop halo), pass 1 delivers it
    for rank in range(world):
        run(rank, deliver=False)
    if symmetric and world == 2:
        # momentum-space data only: the projection reads the extended buffer in place, dataPos is cut out on demand
        ts, loop = run(1, deliver=True, copy_pos_to_host=False)
        assert loop._project_ext and rel_err(loop.dataMom.numpy(), ref_mom[:, :, ts.Tl:]) < TOL_F64
        pos = loop.dataPos_interior().cpu().numpy().reshape(ref.shape[0], 16, 2, ts.Tl, ts.V3h)
        assert rel_err(pos, ref.reshape(ref.shape[0], 16, 2, Lg[3], ts.V3h)[:, :, :, ts.Tl:]) < TOL_F64
    pos_parts, mom_parts = [], []
    for rank in range(world):
        ts, loop = run(rank, deliver=True)
        if symmetric:
            assert loop.tsplit_halo_sides == 1  # every minus-t loop has its plus partner: eigenvector halos travel one way
        pos_parts.append(loop.dataPos.numpy().reshape(ref.shape[0], 16, 2, ts.Tl, ts.V3h))
        mom_parts.append(loop.dataMom.numpy())
    got = np.concatenate(pos_parts, axis=3).reshape(ref.shape)
    assert rel_err(got, ref) < TOL_F64
    assert rel_err(np.concatenate(mom_parts, axis=-1), ref_mom) < TOL_F64
