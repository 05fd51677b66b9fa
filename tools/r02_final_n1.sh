#!/bin/bash
# round-2 single-GPU evidence: GPU tests, the default bench (as the driver runs it), the north_star dictionary,
# ncu launch list of the same command, ncu --set full of the step01 kernels and of K2 / K3, compute-sanitizer memcheck
set -x
python -m pytest tests -q -m gpu > gpurun_out/r02_gputests.log 2>&1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err
python bench.py --dico 2_12 --steps 10 --warmup 3 --no-cpu --no-configs > gpurun_out/r02_bench_n1_dico212.json 2> gpurun_out/r02_bench_n1_dico212.err
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-parity --no-configs > gpurun_out/r02_plain_small.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-parity --no-configs > gpurun_out/r02_ncu_launches.log 2>&1
python tools/step01_probe.py > gpurun_out/r02_plain_step01.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'stream_kernel' -s 6 -c 3 -f -o gpurun_out/r02_step01_stream python tools/step01_probe.py > gpurun_out/r02_ncu_step01.log 2>&1
python tools/profile_one.py 3FWHM > gpurun_out/r02_plain_p1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'fsf_correlate|spectral_glr|local_extrema3' -s 3 -c 3 -f -o gpurun_out/r02_k1_k2_k3 python tools/profile_one.py 3FWHM > gpurun_out/r02_ncu_p1.log 2>&1
