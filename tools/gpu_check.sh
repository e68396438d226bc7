#!/bin/bash
# GPU box helper: parity tests, then the three bench workloads (kernel-only lines).  usage: tools/gpu_check.sh [tag]
tag=${1:-chk}
python -m pytest tests -m gpu -x -q 2>&1 | tail -1
for w in rollout mug reach; do
  python bench.py --workload $w --no-cpu-baseline --no-e2e --no-extra --steps 100 --warmup 5 > gpurun_out/${tag}_$w.json 2> gpurun_out/${tag}_$w.err || { echo "$w FAILED"; tail -3 gpurun_out/${tag}_$w.err; continue; }
  python -c "
import json; d=json.load(open('gpurun_out/${tag}_$w.json')); k=d['config']['kernel']; print('$w', round(d['value']), round(d['ms_per_step'],4), 'ncon', round(d['roofline']['mean_ncon'],2), 'crf', round(d['config']['contact_rich_frac'],3), 'issue-kernel-ms', round(d['roofline']['kernel_ms'],4), 'iters', round(d['roofline']['mean_newton_iters'],3), 'arena', k.get('lite',k)['arena_bytes'], 'wpb', k.get('lite',k)['warps_per_block'])"
done
