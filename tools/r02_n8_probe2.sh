#!/bin/bash
# N=8: paced SM copy kernel for the gather (few blocks) against the copy-engine default
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --no-parity --no-configs "$@"; }
for b in 4 8 16 32; do
  OGN_SCATTER_KERNEL=1 OGN_SCATTER_BLOCKS=$b run > gpurun_out/r02m_b$b.json 2> gpurun_out/r02m_b$b.err
done
python - <<'PY'
import json
for n in ('b4','b8','b16','b32'):
    try:
        d=json.loads(open('gpurun_out/r02m_%s.json'%n).read().strip().splitlines()[-1])
        print(n, round(d['ms_per_step'],3), 'span', round(d['step05_span_ms'],3), [ (r.get('peer_scatter'), r.get('k1_fsf_correlate'), r.get('step05_span')) for r in d['per_rank_stage_ms'][:3]])
    except Exception as e: print(n,'ERR',e)
PY
