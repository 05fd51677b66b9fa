#!/bin/bash
# final single-GPU lines of round 2 (after the K2 FFMA2 change) + memcheck of the smoke path
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err
python bench.py --dico 2_12 --steps 10 --warmup 3 --no-cpu --no-configs > gpurun_out/r02_bench_n1_dico212.json 2> gpurun_out/r02_bench_n1_dico212.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_plain.log 2>&1 &&
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_sanitizer_memcheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_stdout.log 2>&1
tail -3 gpurun_out/r02_sanitizer_memcheck.log; tail -2 gpurun_out/r02_sanitizer_stdout.log
