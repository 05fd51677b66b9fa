#!/bin/bash
# round-2 evidence on 2 GPUs: the whole -m gpu suite (includes tools/check_sharded.py under torchrun) and the N = 2 bench
python -m pytest tests -q -m gpu > gpurun_out/r02_gputests_2gpu.log 2>&1
cp gpurun_out/check_sharded_n2.log gpurun_out/r02_check_sharded_n2.log 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -3 gpurun_out/r02_gputests_2gpu.log; tail -2 gpurun_out/r02_check_sharded_n2.log; tail -2 gpurun_out/r02_bench_n2.err | cut -c1-200
