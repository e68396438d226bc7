#!/bin/bash
# Round-end evidence on the GPU box: tests, smoke, the driver-shaped bench lines, the ncu launch list and one full capture per workload.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err || tail -3 gpurun_out/r2_bench_default.err
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
python bench.py --workload mug --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_mug.json 2> gpurun_out/r2_bench_mug.err
python bench.py --workload reach --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_reach.json 2> gpurun_out/r2_bench_reach.err
python bench.py --dtype f64 --steps 20 --warmup 3 --settle 200 --no-cpu-baseline --no-extra > gpurun_out/r2_bench_f64.json 2> gpurun_out/r2_bench_f64.err
cmd="python bench.py --settle 100 --steps 20 --warmup 5 --no-extra --no-cpu-baseline --no-e2e"
$cmd > gpurun_out/r2_launch_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv $cmd > gpurun_out/r2_launch_ncu.log 2>&1
bash tools/prof.sh rollout r2_prof_rollout
bash tools/prof.sh mug r2_prof_mug
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches.csv
