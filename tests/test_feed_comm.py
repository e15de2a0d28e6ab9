"""The streamed eigenvector feed (mugiq_b200_loop_feed_*) and the library communicator (mugiq_b200_comm_*,
mugiq_b200_allreduce*, loop_plan_accumulate_allreduce) against the CPU oracle.  Single-GPU cases use a one-rank
communicator (the chunked time-slice launches and the grouped NCCL calls still run); the two-rank cases need two GPUs."""
import ctypes as C
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import rel_err, TOL_F64, ROOT
from mugiq_b200 import synth, _lib


def test_feed_and_comm_arguments_are_validated_without_a_gpu():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.mugiq_b200_loop_feed_create(C.byref(h), None, None, 4, 2, 0, 0, None) == -1
    assert b"NULL" in lib.mugiq_b200_last_error()
    assert lib.mugiq_b200_loop_feed_push_host(None, None, None, 1) == -1
    assert lib.mugiq_b200_loop_feed_finish(None, None) == -1
    assert lib.mugiq_b200_loop_feed_destroy(None) == 0
    assert lib.mugiq_b200_comm_create(C.byref(h), None, 0, 1) == -1
    assert lib.mugiq_b200_comm_create(C.byref(h), C.c_char_p(b"x" * 128), 2, 2) == -1
    assert b"rank/size" in lib.mugiq_b200_last_error()
    assert lib.mugiq_b200_allreduce(None, 4, 8, None, None) == -1
    assert lib.mugiq_b200_allreduce_pos(None, None, 0, 0, -1, None, None, None) == -1
    assert lib.mugiq_b200_comm_destroy(None) == 0


L = (8, 4, 4, 8)
ENTRIES = [(0, 1, 1, 2), (0, 0, 1, 2), (3, 1, 1, 1), (3, 0, 1, 1), (1, 1, 2, 2)]


def _inputs(nEv, seed=61):
    return synth.random_evecs_np(L, nEv, seed=seed), synth.sigmas(nEv), synth.random_gauge(L, seed=seed)


@pytest.mark.gpu
def test_feed_push_host_matches_oracle(oracle):
    """Pinned host eigenvectors through the staging ring (batches of 3 out of 7, so the last batch is short), twice on the
    same feed, then on a rebuilt plan (set_plan)."""
    from mugiq_b200 import ops
    nEv = 7
    ev, sig, U = _inputs(nEv)
    ref = oracle.compute_loop(ev, sig, U, ENTRIES, L)
    gd = ops.gauge_upload(U, L)
    plan = ops.LoopPlan(gd, ENTRIES, L)
    pos = torch.full(ref.shape, 3.0, dtype=torch.complex128, device="cuda")
    ev_h = torch.from_numpy(ev).pin_memory()
    feed = ops.LoopFeed(plan, pos, batch=3, nbuf=2)
    for _ in range(2):
        feed.push_host([ev_h[n] for n in range(nEv)], sig)
        assert feed.finish() == nEv
        plan.finalize(pos)
        assert rel_err(pos.cpu().numpy(), ref) < TOL_F64
        pos.fill_(-1.0)
    plan2 = ops.LoopPlan(gd, ENTRIES, L)
    plan.close()
    pos2 = torch.zeros_like(pos)
    feed.set_plan(plan2, pos2)
    feed.push_host([ev_h[n] for n in range(nEv)], sig)   # separate host tensors of one allocation: merged copies
    feed.finish()
    plan2.finalize(pos2)
    assert rel_err(pos2.cpu().numpy(), ref) < TOL_F64
    feed.close()
    plan2.close()


@pytest.mark.gpu
@pytest.mark.parametrize("order", [0, 2, 4])
def test_feed_device_producer(oracle, order):
    """A device-side producer (here: torch copies on a side stream, standing in for QUDA's prolongator) fills the staging
    fields the feed hands out, in the canonical order or in a QUDA native order."""
    from mugiq_b200 import ops
    from oracle import ref_kernels as rk
    nEv = 5
    ev, sig, U = _inputs(nEv, seed=62)
    ref = oracle.compute_loop(ev, sig, U, ENTRIES, L)
    gd = ops.gauge_upload(U, L)
    plan = ops.LoopPlan(gd, ENTRIES, L)
    pos = torch.zeros(ref.shape, dtype=torch.complex128, device="cuda")
    src = [torch.from_numpy(ev[n]).cuda() for n in range(nEv)]
    if order:
        src = [rk.site_to_quda(v, order) for v in src]
    feed = ops.LoopFeed(plan, pos, batch=2, nbuf=2, order=order)
    side = torch.cuda.Stream()
    V4 = ev.shape[1]

    class Raw:  # a device pointer as a torch tensor
        def __init__(self, ptr):
            self.__cuda_array_interface__ = {"shape": (V4 * 12 * 2,), "typestr": "<f8", "data": (ptr, False), "version": 2}

    torch.cuda.synchronize()
    for n0 in range(0, nEv, 2):
        nb = min(2, nEv - n0)
        slots = feed.acquire(nb, side)
        with torch.cuda.stream(side):
            for i in range(nb):
                torch.as_tensor(Raw(slots[i]), device="cuda").copy_(torch.view_as_real(src[n0 + i]).reshape(-1), non_blocking=True)
        feed.commit(sig[n0:n0 + nb], side)
    assert feed.finish() == nEv
    plan.finalize(pos)
    assert rel_err(pos.cpu().numpy(), ref) < TOL_F64
    feed.close()
    plan.close()


@pytest.mark.gpu
def test_feed_rejects_wrong_call_order():
    from mugiq_b200 import ops
    U = synth.random_gauge(L, seed=3)
    plan = ops.LoopPlan(ops.gauge_upload(U, L), ENTRIES[:1], L)
    pos = torch.zeros((3, 16, int(np.prod(L))), dtype=torch.complex128, device="cuda")
    feed = ops.LoopFeed(plan, pos, batch=2)
    with pytest.raises(_lib.MugiqB200Error, match="no batch was acquired"):
        feed.commit([1.0])
    feed.acquire(2)
    with pytest.raises(_lib.MugiqB200Error, match="not committed"):
        feed.acquire(1)
    with pytest.raises(_lib.MugiqB200Error, match="not in"):
        ops.LoopFeed(plan, pos, batch=2).acquire(3)
    with pytest.raises(_lib.MugiqB200Error, match="not committed"):
        feed.finish()
    feed.close()
    plan.close()


def _one_rank_comm():
    lib = _lib.load()
    ident = C.create_string_buffer(_lib.COMM_ID_BYTES)
    _lib.check(lib.mugiq_b200_comm_unique_id(ident))
    h = C.c_void_p()
    _lib.check(lib.mugiq_b200_comm_create(C.byref(h), ident, 0, 1))
    return h


@pytest.mark.gpu
@pytest.mark.parametrize("nchunks", [1, 3, 8, 100])
def test_chunked_accumulate_allreduce_on_one_rank(oracle, nchunks):
    """loop_plan_accumulate_allreduce with a one-rank communicator: the time-slice chunks of the kernels and the grouped
    all-reduces of their runs must reproduce the unchunked result (the sum over one rank is the identity)."""
    from mugiq_b200 import ops
    lib = _lib.load()
    nEv = 5
    ev, sig, U = _inputs(nEv, seed=63)
    ref = oracle.compute_loop(ev, sig, U, ENTRIES, L)
    gd = ops.gauge_upload(U, L)
    plan = ops.LoopPlan(gd, ENTRIES, L)
    evd = [torch.from_numpy(ev[n]).cuda() for n in range(nEv)]
    pos = torch.full(ref.shape, 2.0, dtype=torch.complex128, device="cuda")
    comm = _one_rank_comm()
    info = (C.c_int(), C.c_int(), C.c_int())
    _lib.check(lib.mugiq_b200_comm_info(comm, *[C.byref(x) for x in info]))
    assert (info[0].value, info[1].value) == (0, 1) and info[2].value >= 21800   # NCCL >= 2.18

    class Wrap:
        _h = comm

    plan.accumulate(pos, evd[:2], sig[:2], accumulate=False)
    plan.accumulate_allreduce(pos, evd[2:], Wrap, sigma=sig[2:], accumulate=True, nchunks=nchunks)
    plan.finalize(pos)
    assert rel_err(pos.cpu().numpy(), ref) < TOL_F64
    # the plain entry points on the same communicator
    x = torch.arange(10, dtype=torch.float64, device="cuda")
    _lib.check(lib.mugiq_b200_allreduce(x.data_ptr(), 10, 8, comm, None))
    slots = (C.c_int * 2)(0, 3)
    g = _lib.make_geom(L, 8)
    _lib.check(lib.mugiq_b200_allreduce_pos(pos.data_ptr(), slots, 2, 2, 6, C.byref(g), comm, None))
    torch.cuda.synchronize()
    assert torch.equal(x.cpu(), torch.arange(10, dtype=torch.float64)) and rel_err(pos.cpu().numpy(), ref) < TOL_F64
    _lib.check(lib.mugiq_b200_comm_destroy(comm))
    plan.close()


WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from mugiq_b200 import synth, ops
from mugiq_b200.loop import Loop_Mugiq, Eigsolve
from mugiq_b200.params import MugiqLoopParam, momenta_up_to
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
L = (8, 4, 4, 8); nEv = 9
ev = synth.random_evecs_np(L, nEv, seed=64); sig = synth.sigmas(nEv); U = synth.random_gauge(L, seed=64)
prm = MugiqLoopParam(gauge=[U[mu] for mu in range(4)])
prm.set_displacements("+x:1,2;-x:1,2;+t:1;-t:1;+y:2")
prm.set_momenta(momenta_up_to(1))
lo, hi = rank * nEv // world, (rank + 1) * nEv // world
comm = ops.Comm(dist.group.WORLD)
out = {{}}
for mode, kw in [("overlapped", dict(comm=comm, reduce_pos=True, allreduce_chunks=4)), ("after", dict(comm=comm, reduce_pos=True, allreduce_chunks=1)),
                 ("torch", dict(reduce_pos=True)), ("mom_only", dict(comm=comm, reduce_pos=False)),
                 ("host_streamed", dict(comm=comm, reduce_pos=True, host=True)),
                 ("peer_dma", dict(comm=comm, reduce_pos=True, allreduce_chunks=4, peer_reduce=True))]:
    host = kw.pop("host", False)
    vecs = [torch.from_numpy(ev[n]).pin_memory() if host else torch.from_numpy(ev[n]).cuda() for n in range(lo, hi)]
    loop = Loop_Mugiq(prm, Eigsolve(vecs, sig[lo:hi], L), device=torch.device("cuda", rank), group=dist.group.WORLD, evec_batch=2,
                      stream_batch=2, copy_pos_to_host=False, **kw)
    loop.computeCoarseLoop()
    if mode == "peer_dma":  # a second step on the same mapped buffers (the staging halves alternate)
        loop.MomProjDone = False
        loop.computeCoarseLoop()
    out[mode + "_pos"] = loop.dataPos_d.cpu().numpy()
    out[mode + "_mom"] = loop.dataMom.numpy()
    loop.close_peer_reduce()
np.savez({out!r} + str(rank) + ".npz", **out)
torch.cuda.synchronize()
comm.close()
dist.destroy_process_group()
'''


@pytest.mark.gpu
def test_eigenvector_shards_over_the_library_communicator(oracle, tmp_path):
    """Two ranks, one per GPU: every form of the cross-rank sum (chunked all-reduce overlapped with the kernels, one
    all-reduce after them, torch.distributed, projected buffer only, host-streamed eigenvectors, chunks moved by the copy
    engines over peer-mapped buffers) against the oracle."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL does not put two ranks on one device)")
    from oracle import numpy_check as npc
    from mugiq_b200.params import momenta_up_to
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, out=str(tmp_path / "rank")))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    for p in procs:
        out, err = p.communicate(timeout=600)
        assert p.returncode == 0, err[-3000:]
    Lx, nEv = (8, 4, 4, 8), 9
    ev = synth.random_evecs_np(Lx, nEv, seed=64)
    sig = synth.sigmas(nEv)
    U = synth.random_gauge(Lx, seed=64)
    entries = [(0, 1, 1, 2), (0, 0, 1, 2), (3, 1, 1, 1), (3, 0, 1, 1), (1, 1, 2, 2)]
    ref = oracle.compute_loop(ev, sig, U, entries, Lx)
    ref_mom = npc.momentum_projection(ref, momenta_up_to(1), -1, Lx)
    for rank in range(2):
        z = np.load(tmp_path / f"rank{rank}.npz")
        for mode in ("overlapped", "after", "torch", "host_streamed", "peer_dma"):
            assert rel_err(z[mode + "_pos"], ref) < TOL_F64, (rank, mode)
            assert rel_err(z[mode + "_mom"], ref_mom) < TOL_F64, (rank, mode)
        assert rel_err(z["mom_only_mom"], ref_mom) < TOL_F64
