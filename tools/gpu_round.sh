#!/bin/bash
# One GPU-box round (~6 GPU-minutes): parity tests, smoke, bench (both arms, both kernel paths), complete ncu launch list, DRAM traffic per launch,
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
# every launch of one bench command (prepare ~80 launches, fp64 peak probe, warm-up, timed steps, e2e): -c 400 covers it
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --batch 592 --horizon 6 --no-cpu-baseline > $O/ncu_list_$TAG.log 2>&1
grep -c uniform $O/launches_$TAG.csv
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
# the other BASELINE.json workloads, the rows around the hot path, single-sequence latency
echo "# name | workload | predictions/s objective+gradient | objective only | kernel ms | prepare ms | fp64 executed frac fwd/bwd" > $O/other_$TAG.txt
for w in C2 C3 C4a "C5 --batch 1184 --horizon 10"; do
  timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); f=d['roofline']['fp64']
print('$w'.split()[0], '|', d['config']['workload'], '|', round(d['value']), '|', round(d['forward_only']['value']), '|', d['kernel_ms'], '|', d['prepare_ms']['steady'], '| %.2f/%.2f' % (f['executed_frac_fwd'], f.get('executed_frac_bwd', 0)))" >> $O/other_$TAG.txt
done
cat $O/other_$TAG.txt
timeout 300 python tools/bench_next_rows.py 2>&1 | grep "^N[123]" > $O/next_rows_$TAG.txt; cat $O/next_rows_$TAG.txt
timeout 120 python tools/latency_single.py > $O/latency_$TAG.txt 2>&1; tail -8 $O/latency_$TAG.txt
ls -la $O | tail -20
