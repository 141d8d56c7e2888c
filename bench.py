#!/usr/bin/env python
"""bench.py — agent-env steps/sec (updates included) of the fused hot path on N B200s.

A "step" of this bench is one pass of the hot path over one batch: every agent of every GPU
advances CHUNK training episodes of a reference-style run `agent.train(env, n_episodes,
n_episodes/10)` (src/bin/taxi.rs:165-166), including the evaluate(100) the reference injects
after every episode with episode % eval_at == 0 (src/agent.rs:107-113).  After n_episodes the
agent is reset (src/bin/taxi.rs:200) and the next run starts, so any --steps/--warmup works.

metric  = agent-env steps/sec, updates included: one unit is one iteration of the loop at
          src/agent.rs:86-106 (env.step + get_action + update).  Steps spent inside the
          injected evaluate() calls are executed but NOT counted (reported as eval_steps).
value   = whole-job units / device time (CUDA events, max over ranks), outputs left in HBM.
e2e     = the same through the C ABI with HOST buffers: the per-agent episode records
          (reward_history / episode_length of agent.rs:117) and the per-episode sums are
          copied to pinned host memory inside the timed region.

Default workload = BASELINE.json configs[1] ("c2"): FrozenLake 8x8 slippery, Sarsa(lambda),
eps-greedy, Basic, 1 048 576 agents per GPU (weak scaling), f32.  `--workload c4` is the Taxi
Q-learning target configuration (2 097 152 agents per GPU = 16 M over 8 GPUs).

`--impl reference` times the CPU oracle (the reference cannot be compiled here: no Rust
toolchain) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


WORKLOADS = {
    "c1": dict(desc="Blackjack one-step Q-learning, eps-greedy, Basic", env=0, agent=0, selector=0, policy=0, target=1,
               agents_per_gpu=1 << 22, n_episodes=1000, chunk=100),
    "c2": dict(desc="FrozenLake 8x8 slippery, Sarsa(lambda) eligibility traces, eps-greedy, Basic", env=1, agent=1,
               selector=0, policy=0, target=0, agents_per_gpu=1 << 20, n_episodes=1000, chunk=100, slippery=True),
    "c3": dict(desc="CliffWalking Expected Sarsa, Double policy, UCB", env=2, agent=0, selector=1, policy=1, target=2,
               agents_per_gpu=1 << 22, n_episodes=200, chunk=20),
    "c4": dict(desc="Taxi one-step Q-learning, eps-greedy, Basic", env=3, agent=0, selector=0, policy=0, target=1,
               agents_per_gpu=1 << 21, n_episodes=1000, chunk=100),
}
# not a BASELINE config: the crate's Dyna bin (src/bin/cliffwalking_model.rs), "next" row N4 of SURVEY.md §8(f)
WORKLOADS["dyna"] = dict(desc="CliffWalking one-step Dyna-Q (InternalModelAgent, RandomModel, 10 planning steps), eps-greedy, Basic",
                         env=2, agent=0, selector=0, policy=0, target=1, planning_steps=10, agents_per_gpu=1 << 20, n_episodes=200,
                         chunk=20)
# C5, the full sweep: 4 envs x {Sarsa, Q, Expected Sarsa one-step; Sarsa(lambda), Q(lambda)} x {eps-greedy, UCB} x
# {Basic, Double} = 80 cells, every cell an engine of its own on every GPU, one metric gather per cell per step.
WORKLOADS["c5"] = dict(desc="full sweep: 4 envs x 5 update rules x {eps-greedy, UCB} x {Basic, Double} = 80 cells", cells=[
    dict(env=env, agent=agent, target=target, selector=sel, policy=pol, slippery=True)
    for env in (0, 1, 2, 3) for (agent, target) in ((0, 0), (0, 1), (0, 2), (1, 0), (1, 1)) for sel in (0, 1) for pol in (0, 1)],
    agents_per_gpu=8192, n_episodes=1000, chunk=100)
A_OF_ENV = {0: 2, 1: 4, 2: 4, 3: 6}


def algorithmic_bytes(w, real_size, train_steps, trace_rows):
    """bytes the algorithm must move per agent-step between the table store and the SM.
    one-step Basic: read Q[s'][0..A) + read Q[s][a] + write Q[s][a] + 2 B packed transition = (A+2)*R + 2;
    Double: (2A+3)*R + 2; UCB adds 4A (counts row) + 8 (count RMW); traces add 4*R*A per swept row
    (read e, read Q, write Q, write e)."""
    A, R = A_OF_ENV[w["env"]], real_size
    per_step = ((2 * A + 3) * R + 2) if w["policy"] else ((A + 2) * R + 2)
    if w["selector"]:
        per_step += 4 * A + 8
    k = w.get("planning_steps", 0)
    if k:   # Dyna: membership word + k replays, each one 8-byte model entry and one more update
        per_step = per_step * (1 + k) + 8 * k + 4
    return train_steps * per_step + trace_rows * 4 * R * A


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, power, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=float(max(power)))
        return out


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(workload, agents_per_gpu, dtype):
    """dram__bytes_read.sum + dram__bytes_write.sum per k_run launch from the committed ncu capture of this exact
    configuration (profiles/traffic.json); None when the run's configuration was not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(workload)
        if t and t["agents_per_gpu"] == agents_per_gpu and t["dtype"] == dtype:
            return t["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def workload_hyper(w):
    import parity as P
    return P.hyper(w["n_episodes"], slippery=w.get("slippery", False), planning_steps=w.get("planning_steps", 0))


def combo(w, real):
    return dict(env=w["env"], agent=w["agent"], selector=w["selector"], policy=w["policy"], target=w["target"], real=real)


# ------------------------------------------------------------------------------------------------ CPU legs
def cpu_sample(w, real, n_agents, n_threads, chunk_begin, chunk_end, sessions=None):
    """Advance `n_agents` oracle sessions by episodes [chunk_begin, chunk_end) on `n_threads` host threads.
    Returns (train_steps, seconds, sessions)."""
    import parity as P
    from oracle import oracle_py as O
    from concurrent.futures import ThreadPoolExecutor
    h = workload_hyper(w)
    cfg = P.oracle_config(combo(w, real), h)
    if sessions is None:
        sessions = [O.Session(cfg, i) for i in range(n_agents)]
    L = O.lib()
    eval_at = max(1, w["n_episodes"] // 10)
    lens = [np.zeros(chunk_end - chunk_begin, np.uint64) for _ in sessions]

    def work(tid):
        for i in range(tid, len(sessions), n_threads):
            rc = L.oracle_train(sessions[i].h, chunk_begin, chunk_end, eval_at, None, O._p(lens[i]), None, None)   # ctypes drops the GIL
            assert rc == 0
    t0 = time.perf_counter()
    if n_threads == 1:
        work(0)
    else:
        with ThreadPoolExecutor(n_threads) as ex:
            list(ex.map(work, range(n_threads)))
    dt = time.perf_counter() - t0
    return int(sum(int(l.sum()) for l in lens)), dt, sessions


def run_reference_arm(args, w, real):
    """`--impl reference`: the reference's CPU implementation of the path.  The Rust crate cannot be built in this
    image, so the C++ oracle (oracle/, a line-by-line port with hash-map tables and the per-step allocations the
    reference makes) stands in, one independent agent stream per host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cells = [dict(w, **c) for c in w["cells"]] if "cells" in w else [w]
    per_thread = 64 if len(cells) == 1 else 1   # agents per host thread (per cell): each reference step is a few hundred ms of CPU work
    n_agents = cores * per_thread
    chunk, n_ep = w["chunk"], w["n_episodes"]
    sessions = [None] * len(cells)
    k = 0
    times, steps = [], []
    for it in range(args.warmup + args.steps):
        c = k % (n_ep // chunk)
        ts, dt = 0, 0.0
        for ci, cell in enumerate(cells):
            if c == 0 and k > 0:
                for s in sessions[ci]:
                    s.agent_reset()
            ts_, dt_, sessions[ci] = cpu_sample(cell, real, n_agents, cores, c * chunk, (c + 1) * chunk, sessions[ci])
            ts += ts_; dt += dt_
        k += 1
        if it >= args.warmup:
            times.append(dt); steps.append(ts)
    total_t, total_s = sum(times), sum(steps)
    v = total_s / total_t
    line = {
        "impl": "reference", "metric": "agent-env steps/sec (updates incl.)", "value": v, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if real else "f32", "data": "synthetic",
        "config": {"workload": args.workload + ": " + w["desc"], "agents": n_agents * len(cells), "episodes_per_step": chunk,
                   "n_episodes": n_ep, "eval_at": n_ep // 10},
        "cpu_baseline": {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d agents (%d per host thread%s) x %d-episode chunks of the same run; C++ oracle, hash-map tables"
                                   % (n_agents * len(cells), per_thread, " per cell" if len(cells) > 1 else "", chunk)},
        "e2e": {"value": v, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--agents-per-gpu", type=int, default=0)
    ap.add_argument("--real", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--store", type=int, default=0, help="table store: 0 auto, 1 HBM, 2 shared memory")
    ap.add_argument("--cell-streams", type=int, default=8,
                    help="multi-cell workloads (c5): host threads / CUDA streams the cells' engines are spread over, so that "
                         "launches too small to fill the GPU overlap (1 = back to back on one stream)")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.agents_per_gpu:
        w["agents_per_gpu"] = args.agents_per_gpu
    real = 0 if args.real == "f32" else 1
    real_size = 4 if real == 0 else 8
    if args.warmup < 3 and args.impl == "ours":
        print("note: --warmup %d < 3 (timing rules ask for >= 3)" % args.warmup, file=sys.stderr)

    if args.impl == "reference":
        run_reference_arm(args, w, real)
        return

    import torch
    import torch.distributed as dist
    import parity as P
    rlb = importlib.import_module("rl-rust_b200")
    sh = importlib.import_module("rl-rust_b200.sharding")
    if not torch.cuda.is_available() or rlb.abi.lib.rlb_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path (use --impl reference for the CPU arm)")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    N = w["agents_per_gpu"]
    chunk, n_ep = w["chunk"], w["n_episodes"]
    eval_at = max(1, n_ep // 10)
    chunks_per_run = n_ep // chunk
    # one engine per cell (N agents each on this GPU); every workload but C5 is a single cell
    cells = [dict(w, **c) for c in w["cells"]] if "cells" in w else [w]
    stream = torch.cuda.current_stream()
    n_groups = max(1, min(args.cell_streams, len(cells))) if len(cells) > 1 else 1
    group_streams = [stream] if n_groups == 1 else [torch.cuda.Stream() for _ in range(n_groups)]
    engines = []
    for ci, cell in enumerate(cells):
        e_ = P.make_engine(combo(cell, real), workload_hyper(cell), N, first_agent_id=sh.shard(rank, N), device=local_rank,
                           store_kind=args.store)
        e_.set_stream(group_streams[ci % n_groups].cuda_stream)
        engines.append(e_)
    pool = None
    if n_groups > 1:   # the C ABI's train call returns when its launch is done: one host thread per stream keeps them all busy
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(n_groups)
    eng = engines[0]
    table_bytes = sum(N * e_.S * e_.T * e_.A * real_size for e_ in engines)   # every env's rows are unpadded (Taxi: 24-byte f32 rows)

    sums_dev = [torch.zeros((chunk, 4), dtype=torch.float64, device="cuda") for _ in cells]
    rec_dtype_size = 16 if real == 0 else 32
    state = {"k": 0}
    acc = {"train_steps": 0, "eval_steps": 0, "kernel_ms": 0.0, "launches": 0, "trace_rows": 0, "alg_bytes": 0}

    def run_cell(ci, k, c, host_out):
        e_ = engines[ci]
        if c == 0 and k > 0:
            e_.agent_reset()                                    # src/bin/taxi.rs:200
        if host_out is None:
            return e_.train((c + 1) * chunk, eval_at, ep_begin=c * chunk, sums_out=sums_dev[ci])
        return e_.train((c + 1) * chunk, eval_at, ep_begin=c * chunk, sums_out=host_out[0][ci], episodes_out=host_out[1][ci])

    def step(host_out=None, count=True):
        k = state["k"]
        c = k % chunks_per_run
        if pool is None:
            results = None
        else:                                                   # group g drives cells g, g + n_groups, ... in order on its stream
            def run_group(g):
                torch.cuda.set_device(local_rank)
                return [(ci, run_cell(ci, k, c, host_out)) for ci in range(g, len(cells), n_groups)]
            results = dict(x for grp in pool.map(run_group, range(n_groups)) for x in grp)
        for ci, (cell, e_) in enumerate(zip(cells, engines)):
            r = run_cell(ci, k, c, host_out) if results is None else results[ci]
            if world > 1:                                       # the path's one collective: per-episode metrics to rank 0
                if host_out is not None:
                    sums_dev[ci].copy_(host_out[0][ci], non_blocking=True)
                state["curves"] = sh.gather_episode_sums(sums_dev[ci])   # [world, chunk, 4] on rank 0
            if count:
                acc["train_steps"] += r["train_steps"]; acc["eval_steps"] += r["eval_steps"]
                acc["kernel_ms"] += r["kernel_ms"]; acc["launches"] += 2 * r["kernel_launches"]   # each k_run is followed by one k_episode_sums
                acc["trace_rows"] += r["trace_rows"]
                acc["alg_bytes"] += algorithmic_bytes(cell, real_size, r["train_steps"], r["trace_rows"])
        state["k"] = k + 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(count=False)
    # Timing rules: a run that saw a hardware / thermal slowdown is rejected and measured once more (sw_power_cap is
    # kept and noted in `clocks.reasons`).  Every rank must take the same decision, so the flag is max-reduced.
    remeasured = False
    for attempt in range(2):
        acc.update(train_steps=0, eval_steps=0, kernel_ms=0.0, launches=0, trace_rows=0, alg_bytes=0)
        sampler = ClockSampler(local_rank)
        sampler.start()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        barrier()
        t_wall = time.perf_counter() - t_wall0
        clocks = sampler.stop()
        ms = ev0.elapsed_time(ev1)
        slowed = float(any(r in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown") for r in clocks.get("reasons", [])))
        if world > 1:
            flag = torch.tensor([slowed], dtype=torch.float64, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            slowed = float(flag.item())
        if not slowed or attempt == 1:
            break
        remeasured = True
    clocks["remeasured_after_slowdown"] = remeasured
    dev = acc.copy()

    # ---- e2e leg: same steps, host buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        host_sums = [torch.zeros((chunk, 4), dtype=torch.float64).pin_memory() for _ in cells]
        host_eps = [torch.zeros((chunk, N, rec_dtype_size // 4), dtype=torch.int32).pin_memory() for _ in cells]
        acc.update(train_steps=0, eval_steps=0, kernel_ms=0.0, launches=0, trace_rows=0, alg_bytes=0)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step(host_out=(host_sums, host_eps))
        e1.record(stream)
        barrier()
        e2e_ms = e0.elapsed_time(e1)
        e2e = {"ms": e2e_ms, "train_steps": acc["train_steps"],
               "d2h": len(cells) * (chunk * N * rec_dtype_size + chunk * 32 + 64), "h2d": 24 * len(cells)}

    # ---- reduce over ranks: max time, sum of units
    def allreduce(vals, op):
        if world == 1:
            return vals
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return t.tolist()
    tmax = allreduce([ms, e2e["ms"] if e2e else 0.0, dev["kernel_ms"]], dist.ReduceOp.MAX if world > 1 else None)
    tsum = allreduce([dev["train_steps"], dev["eval_steps"], e2e["train_steps"] if e2e else 0, dev["trace_rows"]],
                     dist.ReduceOp.SUM if world > 1 else None)
    if rank == 0:
        ms_max, e2e_ms_max, kernel_ms_max = tmax
        train_steps, eval_steps, e2e_steps, trace_rows = tsum
        value = train_steps / (ms_max * 1e-3)
        peak, peak_src = peaks()
        # dominant kernel = k_run; per-launch duration measured live by the library (CUDA events on the engine's stream)
        n_launch = max(1, dev["launches"] // 2)                   # k_run launches on this rank (each followed by one k_episode_sums)
        alg_bytes = dev["alg_bytes"]
        achieved = alg_bytes / (dev["kernel_ms"] * 1e-3) / 1e9 if dev["kernel_ms"] > 0 else 0.0
        line = {
            "metric": "agent-env steps/sec (updates incl.)", "value": value, "unit": "agent-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.real, "data": "synthetic",
            "config": {"workload": args.workload + ": " + w["desc"], "agents_per_gpu": N * len(cells), "agents_total": N * len(cells) * world, "cells": len(cells),
                       "episodes_per_step": chunk, "n_episodes": n_ep, "eval_at": eval_at,
                       "table_store": {1: "hbm", 2: "shared_memory_groups", 3: "hybrid_smem_q_l2_traces"}[rlb.abi.lib.rlb_engine_store_kind(eng.h)],
                       "eval_steps_executed_not_counted": eval_steps, "env_steps_per_s_incl_eval": (train_steps + eval_steps) / (ms_max * 1e-3),
                       "l2": "inputs larger than L2: %.2f GB of per-agent tables per GPU vs 126 MB L2 (no flush needed)" % (table_bytes / 1e9),
                       "parallelism": "agents sharded by global id, %d per GPU; one NCCL gather of [episodes,4] metrics per step" % N
                                      + ("; the %d cells' engines spread over %d host threads / CUDA streams (their launches overlap, so "
                                         "kernel_share_of_step counts concurrent kernels)" % (len(cells), n_groups) if n_groups > 1 else ""),
                       "wall_s": t_wall},
            "clocks": clocks,
            "gpu_launches": int(dev["launches"]),
            "roofline": {"bound": "hbm", "kernel": "k_run", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic(args.workload, N, args.real) if len(cells) == 1 else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes / n_launch, "launch_ms": dev["kernel_ms"] / n_launch,
                         "kernel_share_of_step": dev["kernel_ms"] / ms if ms > 0 else None},
        }
        if e2e:
            line["e2e"] = {"value": e2e_steps / (e2e_ms_max * 1e-3), "unit": "agent-steps/s",
                           "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                           "note": "C-ABI rlb_agent_train_range with pinned HOST buffers: per-agent episode records + per-episode sums copied "
                                   "device->host every step; the path has no per-step host inputs besides the call's scalar arguments"}
        if world == 1 and not args.no_cpu_baseline:
            cpu_agents = 512 if len(cells) == 1 else 4   # ~10 s of single-thread CPU work
            ts, dt = 0, 0.0
            for cell in cells:
                ts_, dt_, _ = cpu_sample(cell, real, cpu_agents, 1, 0, n_ep)
                ts += ts_; dt += dt_
            line["cpu_baseline"] = {"value": ts / dt, "unit": "agent-steps/s", "cores": 1, "kind": "port",
                                    "sample": "%d agents%s x one full %d-episode run (eval_at %d), C++ oracle single thread, %.1f s"
                                              % (cpu_agents, " per cell" if len(cells) > 1 else "", n_ep, eval_at, dt)}
        print(json.dumps(line), flush=True)
    if pool is not None:
        pool.shutdown()
    for e_ in engines:
        e_.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
