#!/bin/bash
# Debug-build memory-safety check (compute-sanitizer is closed on the GPU pool): builds the library with guard words between the arena's
# arrays (-DUR3E_CANARY) into variants/ here (CPU), then -- on the GPU box -- runs the GPU parity tests against that build.
#   here:        tools/canary_check.sh build
#   GPU box:     tools/canary_check.sh run      (a clobbered guard word prints "arena guard word clobbered" and fails)
set -e
cd "$(dirname "$0")/.."
mkdir -p variants/canary_obj gpurun_out
if [ "$1" = build ]; then
  UR3E_EXTRA_FLAGS="-DUR3E_CANARY" UR3E_OBJ_DIR=$PWD/variants/canary_obj UR3E_LIB_OUT=$PWD/variants/libur3e_b200_canary.so python -c "from ur3e_b200 import build as b; print(b.build())"
else
  UR3E_B200_LIB=$PWD/variants/libur3e_b200_canary.so python -m pytest tests -m gpu -x -q -s 2>&1 | tee gpurun_out/canary.log | tail -5
  if grep -q "guard word clobbered" gpurun_out/canary.log; then echo "CANARY FAILED"; exit 1; else echo "canary build: no guard word clobbered"; fi
fi
