#!/bin/bash
# last 2-GPU check of the round: the sharded-parity test (tools/check_sharded.py under torchrun) and a short N = 2 bench
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_baseline_sizes.py -q -m gpu -k "sharded" > gpurun_out/r02_final_sharded_test.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_final_sharded_test.log
tail -n 3 gpurun_out/r02_final_sharded_test.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_final_bench_n2.json 2> gpurun_out/r02_final_bench_n2.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02_final_bench_n2.json
