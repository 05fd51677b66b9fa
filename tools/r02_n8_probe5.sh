#!/bin/bash
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --no-parity --no-configs "$@"; }
OGN_TIMING_OFFSETS=1 run > gpurun_out/r02u_auto.json 2> gpurun_out/r02u_auto.err
OGN_BENCH_STAGGER_US=290 run > gpurun_out/r02u_290.json 2> gpurun_out/r02u_290.err
OGN_BENCH_STAGGER_US=0 run > gpurun_out/r02u_0.json 2> gpurun_out/r02u_0.err
python - <<'PY'
import json
for n in ('auto','290','0'):
    try:
        d=json.loads(open('gpurun_out/r02u_%s.json'%n).read().strip().splitlines()[-1])
        print(n, round(d['ms_per_step'],3), 'span', round(d['step05_span_ms'],3), [ (r.get('peer_scatter'), r.get('k1_fsf_correlate'), r.get('step05_span')) for r in d['per_rank_stage_ms'][:8:2]])
    except Exception as e: print(n,'ERR',e)
PY
