#!/bin/bash
# step04 / step08 with the O(k) tridiagonal solver (and step08 on the Gram matrix): parity tests, then the timing probe
# for several Krylov sizes (OGN_PCA_KRYLOV / OGN_LINES_KRYLOV; the defaults are 24 / 40)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_lines.py tests/test_gpu_pca.py -x -q -m gpu > gpurun_out/lines_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/lines_tests.log
tail -n 4 gpurun_out/lines_tests.log
timeout 100 python tools/pca_lines_probe.py gpu > gpurun_out/krylov_default.json 2> gpurun_out/krylov_default.err
for m in "12 12" "16 16" "32 24" "48 64"; do
  set -- $m
  OGN_PCA_KRYLOV=$1 OGN_LINES_KRYLOV=$2 timeout 100 python tools/pca_lines_probe.py gpu > gpurun_out/krylov_$1_$2.json 2> gpurun_out/krylov_$1_$2.err
done
OGN_LINES_NO_GRAM=1 timeout 100 python tools/pca_lines_probe.py gpu > gpurun_out/krylov_default_nogram.json 2>&1
for f in gpurun_out/krylov_*.json; do echo $f; tail -n 1 $f; done
