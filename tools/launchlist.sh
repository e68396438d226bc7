#!/bin/bash
# GPU box helper: ncu launch list (gpu__time_duration.sum of every kernel) of a short default-workload bench run.
cmd="python bench.py --settle 100 --steps 20 --warmup 5 --no-extra --no-cpu-baseline --no-e2e"
$cmd > gpurun_out/r2_launch_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv $cmd > gpurun_out/r2_launch_ncu.log 2>&1
tail -2 gpurun_out/r2_launch_ncu.log; ls -la gpurun_out/r2_launches.csv
