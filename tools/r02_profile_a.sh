#!/bin/bash
# round-2 baseline captures of the shipped K2f v2 and step01 kernels (before this round's changes)
set -x
python tools/profile_one.py 2_12 > gpurun_out/r02a_plain_k2f.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:folded_glr -s 1 -c 1 -f -o gpurun_out/r02a_k2f_v2 python tools/profile_one.py 2_12 > gpurun_out/r02a_ncu_k2f.log 2>&1
python tools/step01_probe.py > gpurun_out/r02a_plain_step01.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'dct_|standardise' -s 4 -c 4 -f -o gpurun_out/r02a_step01 python tools/step01_probe.py > gpurun_out/r02a_ncu_step01.log 2>&1
python bench.py --dico 2_12 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02a_bench_212.json 2> gpurun_out/r02a_bench_212.err
nvidia-smi topo -m > gpurun_out/r02a_topo.txt 2>&1
lscpu > gpurun_out/r02a_lscpu.txt 2>&1; free -g >> gpurun_out/r02a_lscpu.txt
