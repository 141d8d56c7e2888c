#!/bin/bash
# r02m (8 GPUs): the default bench line under torchrun as the driver launches it — headline C2 (weak scaling) and the
# sub-records, C4 = 16 777 216 agents sharded over the eight GPUs (the configuration north_star's 1e11 target is stated on).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $O/r02m_gpus.txt
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r02m_bench_8gpu.out 2> $O/r02m_bench_8gpu.err; echo "bench exit $?"
grep '^{"metric' $O/r02m_bench_8gpu.out > $O/r02m_bench_8gpu.json; cut -c1-220 $O/r02m_bench_8gpu.json
grep -h "Init COMPLETE" $O/r02m_bench_8gpu.out $O/r02m_bench_8gpu.err | head -16 > $O/r02m_nccl_init.txt; wc -l $O/r02m_nccl_init.txt
tail -4 $O/r02m_bench_8gpu.err | cut -c1-300
timeout 600 python tools/rl_bins.py taxi -n 200 --n_agents 65536 --gpus 8 --real f32 --tally_games 0 --out $O/r02m_driver_taxi_8gpu.json > $O/r02m_driver.log 2>&1; tail -13 $O/r02m_driver.log | cut -c1-160
