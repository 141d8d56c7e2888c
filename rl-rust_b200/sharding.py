"""Multi-GPU host logic: agents are independent, so the path shards by contiguous ranges of GLOBAL agent id with no
data-path collective; the only exchange is one gather of the per-episode metric sums ([episodes, 4] f64 per rank)
to rank 0 (NCCL over NVLink on GPUs; gloo in the CPU tests).  Works with any initialised torch.distributed backend."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard(rank, agents_per_rank):
    """Global agent ids [first, first + agents_per_rank) owned by `rank` (Philox counters carry the global id)."""
    return rank * agents_per_rank


def gather_episode_sums(local_sums, dst=0):
    """local_sums: [E,4] f64 tensor (CUDA for nccl, CPU for gloo).  Returns [world,E,4] on dst, None elsewhere."""
    rank, n = world()
    if n == 1:
        return local_sums.unsqueeze(0)
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
