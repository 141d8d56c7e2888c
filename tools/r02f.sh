#!/bin/bash
# r02f (2 GPUs): the multi-GPU paths — rlb_comm_* with two devices (single-process group form), driver --gpus 2, and the
# default bench line under torchrun with its sub-records (C4 = 16 777 216 agents SHARDED over the two GPUs, 100 GB of tables each).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $O/r02f_gpus.txt
timeout 600 python -m pytest tests/test_gpu_abi2.py::test_comm_gather_single_process tests/test_driver.py::test_driver_multi_gpu_matches_single_gpu tests/test_gpu_api.py -m gpu -q > $O/r02f_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/r02f_pytest.log
tail -4 $O/r02f_pytest.log
timeout 300 python bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > $O/r02f_ref.json 2>> $O/r02f_err.log; cut -c1-160 $O/r02f_ref.json
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02f_bench_2gpu.out 2> $O/r02f_bench_2gpu.err; echo "bench exit $?"
grep '^{"metric' $O/r02f_bench_2gpu.out > $O/r02f_bench_2gpu.json; cut -c1-220 $O/r02f_bench_2gpu.json
grep -c "NCCL INFO" $O/r02f_bench_2gpu.out $O/r02f_bench_2gpu.err | head; grep -h "nranks\|NVLS\|Connected all" $O/r02f_bench_2gpu.out $O/r02f_bench_2gpu.err | head -8
tail -5 $O/r02f_bench_2gpu.err | cut -c1-300
timeout 600 python tools/rl_bins.py taxi -n 200 --n_agents 4096 --gpus 2 --real f32 --tally_games 0 --out $O/r02f_driver_taxi_2gpu.json > $O/r02f_driver.log 2>&1; tail -3 $O/r02f_driver.log
