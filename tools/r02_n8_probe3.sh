#!/bin/bash
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --no-parity --no-configs "$@"; }
OGN_SCATTER_KERNEL=1 OGN_SCATTER_BLOCKS=32 run > gpurun_out/r02n_early_b32.json 2> gpurun_out/r02n_early_b32.err
OGN_SCATTER_KERNEL=1 OGN_SCATTER_BLOCKS=64 run > gpurun_out/r02n_early_b64.json 2> gpurun_out/r02n_early_b64.err
OGN_BENCH_GATHER_LATE=1 OGN_SCATTER_KERNEL=1 OGN_SCATTER_BLOCKS=32 run > gpurun_out/r02n_late_b32.json 2> gpurun_out/r02n_late_b32.err
run > gpurun_out/r02n_early_dma.json 2> gpurun_out/r02n_early_dma.err
python - <<'PY'
import json
for n in ('early_b32','early_b64','late_b32','early_dma'):
    try:
        d=json.loads(open('gpurun_out/r02n_%s.json'%n).read().strip().splitlines()[-1])
        print(n, round(d['ms_per_step'],3), 'span', round(d['step05_span_ms'],3), [ (r.get('peer_scatter'), r.get('k1_fsf_correlate'), r.get('step05_span')) for r in d['per_rank_stage_ms'][:3]])
    except Exception as e: print(n,'ERR',e)
PY
