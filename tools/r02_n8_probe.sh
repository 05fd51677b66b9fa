#!/bin/bash
# N=8 timeline probes: default step, step without the gather, host-synchronous steps
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --no-parity --no-configs "$@"; }
run > gpurun_out/r02l_default.json 2> gpurun_out/r02l_default.err
OGN_BENCH_NO_GATHER=1 run > gpurun_out/r02l_nogather.json 2> gpurun_out/r02l_nogather.err
OGN_SCATTER_KERNEL=1 run > gpurun_out/r02l_smcopy.json 2> gpurun_out/r02l_smcopy.err
python - <<'PY'
import json
for n in ('default','nogather','smcopy'):
    try:
        d=json.loads(open('gpurun_out/r02l_%s.json'%n).read().strip().splitlines()[-1])
        print(n, round(d['ms_per_step'],3), 'span', round(d['step05_span_ms'],3), [ (r.get('peer_scatter'), r.get('k1_fsf_correlate'), r.get('step05_span')) for r in d['per_rank_stage_ms']])
    except Exception as e: print(n,'ERR',e)
PY
