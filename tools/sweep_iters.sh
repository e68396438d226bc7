#!/bin/bash
# measurement aid: throughput vs Newton iteration cap (float32 rollout workload)
for it in 2 3 4 6 8; do
  python bench.py --steps 100 --warmup 20 --no-cpu-baseline --no-e2e --solver-iters $it 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], d['value'], d['roofline_fp32']['mean_newton_iters'], d['config']['unstable_resets'])" $it
done
