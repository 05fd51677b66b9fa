#!/bin/bash
# Footprint-restricted mosaic path: parity tests, then the timing probe with and without the restriction.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "mosaic or multifield or two_field or default_kernel or fsf_stage or symmetry" > gpurun_out/mf_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/mf_tests.log
tail -5 gpurun_out/mf_tests.log
for nf in 2 4; do
  timeout 120 python tools/multifield_probe.py 3681 320 320 $nf > gpurun_out/mf_probe_nf${nf}.log 2>&1
  OGN_K1_NO_FOOTPRINT=1 timeout 120 python tools/multifield_probe.py 3681 320 320 $nf > gpurun_out/mf_probe_nf${nf}_full.log 2>&1
done
tail -4 gpurun_out/mf_probe_nf*.log
