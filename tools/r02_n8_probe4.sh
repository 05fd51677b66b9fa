#!/bin/bash
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --no-parity --no-configs "$@"; }
run > gpurun_out/r02o_dma_auto.json 2> gpurun_out/r02o_dma_auto.err
OGN_BENCH_STAGGER_US=300 run > gpurun_out/r02o_dma_300.json 2> gpurun_out/r02o_dma_300.err
OGN_SCATTER_KERNEL=1 run > gpurun_out/r02o_sm_auto.json 2> gpurun_out/r02o_sm_auto.err
OGN_SCATTER_KERNEL=1 OGN_BENCH_STAGGER_US=300 run > gpurun_out/r02o_sm_300.json 2> gpurun_out/r02o_sm_300.err
OGN_BENCH_GATHER_EARLY=1 run > gpurun_out/r02o_early_dma.json 2> gpurun_out/r02o_early_dma.err
python - <<'PY'
import json
for n in ('dma_auto','dma_300','sm_auto','sm_300','early_dma'):
    try:
        d=json.loads(open('gpurun_out/r02o_%s.json'%n).read().strip().splitlines()[-1])
        print(n, round(d['ms_per_step'],3), 'span', round(d['step05_span_ms'],3), [ (r.get('peer_scatter'), r.get('fsf_prep'), r.get('den_table'), r.get('k1_fsf_correlate'), r.get('step05_span')) for r in d['per_rank_stage_ms'][:8:2]])
    except Exception as e: print(n,'ERR',e)
PY
