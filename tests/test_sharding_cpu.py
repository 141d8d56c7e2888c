"""CPU test of the N>1 host path with gloo, world_size 2: each rank runs its shard of agents (through the oracle, as no
GPU is here), rank 0 gathers the per-episode sums, and the combined curves / totals must equal one 2x-sized run —
sharding by global agent id changes nothing."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_PER_RANK, N_EP, EVAL_AT = 6, 12, 4
CASE = dict(env=1, agent=1, selector=0, policy=0, target=0, real=1)


def _sums(o):
    return np.stack([o["len"].sum(0).astype(np.float64), o["ret"].sum(0), o["tdsum"].sum(0), o["tdabs"].sum(0)], 1)


def _worker(rank, world_size, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity as P
    from oracle import oracle_py as O
    sh = importlib.import_module("rl-rust_b200.sharding")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    first = sh.shard(rank, N_PER_RANK)
    o = O.batch_train(P.oracle_config(CASE, P.hyper(N_EP)), first, N_PER_RANK, N_EP, EVAL_AT)
    gathered = sh.gather_episode_sums(torch.from_numpy(_sums(o)))
    t, u = sh.job_totals(10.0 + rank, o["train_steps"])
    if rank == 0:
        q.put((sh.combine_episode_sums(gathered).numpy(), t, u, first))
    else:
        assert gathered is None and first == N_PER_RANK
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_run():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity as P
    from oracle import oracle_py as O
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    combined, t, u, first0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    whole = O.batch_train(P.oracle_config(CASE, P.hyper(N_EP)), 0, 2 * N_PER_RANK, N_EP, EVAL_AT)
    ref = _sums(whole)
    assert first0 == 0
    assert np.array_equal(combined[:, :2], ref[:, :2])                       # lengths and returns: exact
    assert np.allclose(combined[:, 2:], ref[:, 2:], rtol=1e-12, atol=1e-12)  # float sums regrouped across ranks
    assert t == 11.0 and u == whole["train_steps"]                          # time = max over ranks, units = sum
