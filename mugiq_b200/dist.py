"""Multi-GPU partitioning of the loop path: eigenvectors are sharded over ranks (data-parallel
over n, gauge field replicated); each rank accumulates its shard into a full-volume loop buffer and one
allreduce(sum) over NCCL/NVLink finishes the eigenvector sum.  Replaces the reference's host-staged
MPI_Reduce / MPI_Gather / MPI_Bcast (lib/loop_mugiq.cpp:406-424), which reduce over a *spatial* split that
does not exist here."""
import os


def shard_range(nEv, rank, world):
    """Contiguous block of eigenvector indices owned by `rank`: sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(int(nEv), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def allreduce_loop_buffer(buf, group=None):
    """Sum a complex loop buffer over the group in place (NCCL on CUDA tensors, gloo on CPU tensors)."""
    import torch
    import torch.distributed as dist
    dist.all_reduce(torch.view_as_real(buf) if buf.is_complex() else buf, op=dist.ReduceOp.SUM, group=group)
    return buf
