#!/bin/bash
# A/B helper for the GPU box: bench every library variant under variants/*.so (built with UR3E_EXTRA_FLAGS / UR3E_LIB_OUT, see ur3e_b200/build.py).
# usage: tools/ab.sh [bench args]   -> one line per variant: name value ms_per_step
for v in variants/*.so; do
  UR3E_B200_LIB=$PWD/$v python bench.py --no-cpu-baseline --no-e2e --no-extra --steps 100 --warmup 5 "$@" > /tmp/ab.json 2> /tmp/ab.err || { echo "$v FAILED"; tail -3 /tmp/ab.err; continue; }
  python -c "
import json; d=json.load(open('/tmp/ab.json')); print('$v', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['mean_newton_iters'],3))"
done
