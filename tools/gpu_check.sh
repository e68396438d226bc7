#!/bin/bash
# GPU box helper: parity tests (unless NOTEST=1), then the three bench workloads (kernel-only lines).  usage: tools/gpu_check.sh [tag]
tag=${1:-chk}
mkdir -p gpurun_out
if [ -z "$NOTEST" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -${TESTTAIL:-3}; fi
for w in ${WORKLOADS:-rollout mug reach}; do
  python bench.py --workload $w --no-cpu-baseline --no-e2e --no-extra --steps 100 --warmup 5 > gpurun_out/${tag}_$w.json 2> gpurun_out/${tag}_$w.err || { echo "$w FAILED"; tail -3 gpurun_out/${tag}_$w.err; continue; }
  python -c "
import json; d=json.load(open('gpurun_out/${tag}_$w.json')); k=d['config']['kernel']; c=d['config']; print('$w', round(d['value']), 'ms', round(d['ms_per_step'],4), 'ncon', round(c['mean_ncon'],2), 'nefc', round(c['mean_nefc'],1), 'crf', round(c['contact_rich_frac'],3), 'lite/full ms', round(c['kernel_ms_lite_tier'],3), round(c['kernel_ms_full_tier'],3), 'iters', round(d['roofline']['mean_newton_iters'],3), 'episodes', c['episodes'], 'trunc', c['truncations'], 'ovf', c['overflow_steps'], 'arena', k.get('lite',k)['arena_bytes'], 'wpb', k.get('lite',k)['warps_per_block'], 'regs', k.get('lite',k)['regs_per_thread'], k['regs_per_thread'])"
done
