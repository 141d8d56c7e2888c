"""Multi-GPU host logic: agents are independent, so the path shards by contiguous ranges of GLOBAL agent id with no
data-path collective; the only exchange is one gather of the per-episode metric sums ([episodes, 4] f64 per rank) to
rank 0 for the training charts (bin/taxi.rs:170-223).

On GPUs the gather is the library's own: `rlb_comm_gather_episode_sums` (grouped ncclSend / ncclRecv over NVLink,
include/rlb.h).  torch.distributed is plumbing only: it carries the 128-byte NCCL id from rank 0 to the other ranks, the
barriers and the timing reductions of bench.py — and, with the `gloo` backend, stands in for the gather in the CPU
tests of this host logic (there is no NCCL without a GPU)."""
import torch
import torch.distributed as dist

from . import _abi as abi


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard(rank, agents_per_rank):
    """Global agent ids [first, first + agents_per_rank) owned by `rank` (Philox counters carry the global id)."""
    return rank * agents_per_rank


def shard_sizes(agents_total, n_ranks):
    """Strong scaling: `agents_total` agents cut into n_ranks contiguous shards; (first_id, count) per rank."""
    base, extra = divmod(agents_total, n_ranks)
    out, first = [], 0
    for r in range(n_ranks):
        n = base + (1 if r < extra else 0)
        out.append((first, n))
        first += n
    return out


def make_comm(device):
    """One rlb_comm per rank of the initialised torch.distributed job (None for a single process): rank 0 draws the
    NCCL unique id, torch.distributed broadcasts its 128 bytes."""
    rank, n = world()
    if n == 1:
        return None
    uid = [abi.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    return abi.Comm.init_rank(uid[0], n, rank, device)


def gather_episode_sums(local_sums, dst=0, comm=None, out=None, stream=None):
    """local_sums: [E,4] f64 tensor.  Returns [world,E,4] on dst, None elsewhere.
    With `comm` (an rlb_comm, CUDA tensors) the library's NCCL gather is used; without it torch.distributed (gloo in the
    CPU tests)."""
    rank, n = world()
    if n == 1:
        return local_sums.unsqueeze(0)
    if comm is not None:
        if rank == dst and out is None:
            out = torch.empty((n,) + tuple(local_sums.shape), dtype=local_sums.dtype, device=local_sums.device)
        comm.gather_episode_sums(local_sums, out if rank == dst else None, root=dst, stream=stream)
        return out if rank == dst else None
    bufs = [torch.empty_like(local_sums) for _ in range(n)] if rank == dst else None
    dist.gather(local_sums, bufs, dst=dst)
    return torch.stack(bufs) if rank == dst else None


def combine_episode_sums(gathered):
    """[world,E,4] -> [E,4]: per-episode totals over every agent of the job (sum length, return, td, |td|)."""
    return gathered.sum(0)


def job_totals(elapsed_ms, units, device=None):
    """Whole-job accounting the bench contract asks for: time = MAX over ranks, units = SUM over ranks."""
    rank, n = world()
    if n == 1:
        return float(elapsed_ms), float(units)
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())
