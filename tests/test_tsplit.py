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
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(ok)]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_halo_exchange_gloo(tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(int(np.load(tmp_path / f"ok{r}.npy")[0]) == 1 for r in range(world))


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 4])
def test_tsplit_loop_matches_global_oracle(oracle, world, monkeypatch):
    """Each (virtual) rank computes its time-slab with the fused kernels on the extended lattice; the stitched
    position-space buffer and the gathered momentum-space buffer equal the oracle's on the global lattice.  The halos
    are cut from the global field here (one process); the NCCL/gloo exchange itself is covered above."""
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
    pos_parts, mom_parts = [], []
    for rank in range(world):
        ts = TSplit(Lg, rank, world, max_t_disp=2)
        monkeypatch.setattr(ts, "begin_extend", lambda vecs, group=None, device=None, ts=ts: len(vecs))
        monkeypatch.setattr(ts, "finish_extend", lambda n, ts=ts: ts.global_slab(evg[:n], site_dim=1))
        prm = MugiqLoopParam(gauge=[U[mu] for mu in range(4)])
        prm.set_displacements(entries_str)
        prm.set_momenta(mom)
        inner = _interior_of_global(ts, evg, 1).contiguous()
        loop = Loop_Mugiq(prm, Eigsolve(list(inner), sig, ts.L_loc), tsplit=ts, stream_batch=nEv,
                          group=object() if world > 1 else None)
        monkeypatch.setattr(ts, "gather_time", lambda m, group=None: m)
        loop.computeCoarseLoop()
        pos_parts.append(loop.dataPos.numpy().reshape(ref.shape[0], 16, 2, ts.Tl, ts.V3h))
        mom_parts.append(loop.dataMom.numpy())
    got = np.concatenate(pos_parts, axis=3).reshape(ref.shape)
    assert rel_err(got, ref) < TOL_F64
    assert rel_err(np.concatenate(mom_parts, axis=-1), ref_mom) < TOL_F64
