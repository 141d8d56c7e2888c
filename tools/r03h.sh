#!/bin/bash
# r03h (2 GPUs): the multi-GPU paths on the kernels of record — rlb_comm tests, the default bench line under torchrun
# (C4 = 16 777 216 agents sharded), the reference arm as the driver launches it, driver --gpus 2.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_abi2.py tests/test_driver.py -m gpu -q -k "comm or multi_gpu" > $O/r03h_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r03h_pytest.log
tail -3 $O/r03h_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > $O/r03h_ref.out 2>> $O/r03h_err.log; grep '^{' $O/r03h_ref.out | cut -c1-200
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r03h_bench_2gpu.out 2> $O/r03h_bench_2gpu.err; echo "bench exit $?"
grep '^{"metric' $O/r03h_bench_2gpu.out > $O/r03h_bench_2gpu.json; cut -c1-220 $O/r03h_bench_2gpu.json
tail -3 $O/r03h_bench_2gpu.err | cut -c1-300
timeout 300 python tools/rl_bins.py taxi -n 200 --n_agents 4096 --gpus 2 --real f32 --tally_games 0 --out $O/r03h_driver_taxi_2gpu.json > $O/r03h_driver.log 2>&1; tail -2 $O/r03h_driver.log | cut -c1-200
