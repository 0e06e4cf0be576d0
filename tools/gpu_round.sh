#!/bin/bash
# One GPU-box round: parity tests, smoke, bench, ncu launch list + one full capture of the hot kernel.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_$TAG.txt
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu_$TAG.txt
cat gpurun_out/pytest_gpu_$TAG.txt
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ) > gpurun_out/smoke_$TAG.txt
cat gpurun_out/smoke_$TAG.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --steps 2 --warmup 3 --distinct-lengthscales --no-cpu-baseline > gpurun_out/bench_distinct_$TAG.json 2>> gpurun_out/bench_$TAG.err
tail -c 1500 gpurun_out/bench_distinct_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --batch 592 --horizon 6 --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1
tail -3 gpurun_out/ncu_list_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:uniform_ -s 8 -c 2 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 3 --batch 296 --horizon 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -f -o gpurun_out/prof_general_$TAG \
    python bench.py --steps 1 --warmup 3 --batch 296 --horizon 2 --no-cpu-baseline --distinct-lengthscales > gpurun_out/ncu_full_general_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_general_$TAG.log | cut -c1-200
ls -la gpurun_out
