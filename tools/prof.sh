#!/bin/bash
# GPU box helper: one `ncu --set full` capture of the step kernels of a bench workload inside its timed region.
# usage: tools/prof.sh <workload> <tag> [settle]     (captures the consecutive step_kernel launches of one env-step: lite, grasp and generic tier)
w=$1; tag=$2; settle=${3:-}
[ -z "$settle" ] && { case $w in rollout) settle=3000;; mug) settle=1500;; reach) settle=1000;; esac; }
cmd="python bench.py --workload $w --settle $settle --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-extra"
# launches of step_kernel before the timed region: 4 per env-step on main.xml (lite, generic tier on the side stream, grasp tier,
# generic tier for the grasp tier's overflows), 1 on the others
per=4; [ "$w" = reach ] && per=1
skip=$(( (settle + 3) * per + per ))
$cmd > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s $skip -c $per -o gpurun_out/$tag -f $cmd > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
