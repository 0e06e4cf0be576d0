#!/bin/bash
# One GPU-box round: parity tests, smoke, bench (both arms, both kernel paths), ncu launch list, DRAM traffic per launch,
# one full capture of each hot kernel, per-CTA scheduling diagnostic.   Usage (repo root, under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu_$TAG.txt
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/pytest_gpu_$TAG.txt
tail -3 $O/pytest_gpu_$TAG.txt
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 ) > $O/smoke_$TAG.txt
cat $O/smoke_$TAG.txt
timeout 600 python bench.py --steps 3 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err
tail -c 600 $O/bench_$TAG.json; tail -5 $O/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_reference_$TAG.json 2>> $O/bench_$TAG.err
tail -c 400 $O/bench_reference_$TAG.json
timeout 600 python bench.py --steps 2 --warmup 3 --distinct-lengthscales --no-cpu-baseline > $O/bench_distinct_$TAG.json 2>> $O/bench_$TAG.err
GPMPC_DEBUG_CLOCKS=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | tail -3 > $O/cta_life_$TAG.txt
cat $O/cta_life_$TAG.txt | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --batch 592 --horizon 6 --no-cpu-baseline > $O/ncu_list_$TAG.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:uniform_ -s 4 -c 2 --csv \
    --log-file $O/traffic_u_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:rollout_kernel -s 2 -c 1 --csv \
    --log-file $O/traffic_g_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --distinct-lengthscales > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:uniform_ -s 8 -c 2 -f -o $O/prof_$TAG \
    python bench.py --steps 1 --warmup 3 --batch 2368 --horizon 3 --no-cpu-baseline > $O/ncu_full_$TAG.log 2>&1
tail -1 $O/ncu_full_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -f -o $O/prof_general_$TAG \
    python bench.py --steps 1 --warmup 3 --batch 296 --horizon 2 --no-cpu-baseline --distinct-lengthscales > $O/ncu_full_general_$TAG.log 2>&1
tail -1 $O/ncu_full_general_$TAG.log | cut -c1-200
ls -la $O | tail -20
