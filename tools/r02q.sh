#!/bin/bash
# r02q (8 GPUs): BASELINE configs[4] — the full sweep, 80 cells x 102 400 agents per GPU = 65.5 M agent slots on the box,
# every cell's per-episode sums gathered to rank 0 by rlb_comm_gather_episode_sums (NCCL) every step.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --workload c5 --agents-per-gpu 102400 --steps 3 --warmup 3 --no-e2e --cell-streams 8 --sub '' > $O/r02q_c5_8gpu.out 2> $O/r02q_c5_8gpu.err; echo "bench exit $?"
grep '^{"metric' $O/r02q_c5_8gpu.out > $O/r02q_bench_c5_8gpu.json; cut -c1-300 $O/r02q_bench_c5_8gpu.json
tail -4 $O/r02q_c5_8gpu.err | cut -c1-300
