#!/bin/bash
# last single-GPU check of the round: every GPU test, smoke(), the default bench as the driver runs it
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r02_final_gputests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_final_gputests.log
tail -n 4 gpurun_out/r02_final_gputests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_final_smoke.log 2>&1
tail -n 2 gpurun_out/r02_final_smoke.log
timeout 400 python bench.py > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_final_bench_n1.json
