"""Eigenvector sharding + loop-buffer allreduce on CPU with world_size 2 (gloo): each rank contracts its
shard with the oracle, the allreduced buffer must equal the unsharded result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mugiq_b200 import synth
from mugiq_b200.dist import shard_range, allreduce_loop_buffer

L = (4, 4, 4, 4)
NEV = 5
ENTRIES = [(0, 1, 1, 1), (3, 0, 1, 2)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    orc.set_num_threads(1)
    ev = synth.random_evecs_np(L, NEV, seed=21)
    sig = synth.sigmas(NEV)
    U = synth.random_gauge(L, seed=21)
    lo, hi = shard_range(NEV, rank, world)
    part = orc.compute_loop(ev[lo:hi], sig[lo:hi], U, ENTRIES, L)
    buf = torch.from_numpy(part)
    allreduce_loop_buffer(buf)
    if rank == 0:
        np.save(out_path, buf.numpy())
    dist.destroy_process_group()


def test_sharded_sum_equals_full(tmp_path, oracle):
    out = str(tmp_path / "sum.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    ev = synth.random_evecs_np(L, NEV, seed=21)
    full = oracle.compute_loop(ev, synth.sigmas(NEV), synth.random_gauge(L, seed=21), ENTRIES, L)
    assert np.abs(got - full).max() / np.abs(full).max() < 1e-13
