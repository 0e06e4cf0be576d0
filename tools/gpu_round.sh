#!/bin/bash
# One GPU-box round (~7 GPU-minutes): parity tests, smoke, bench (both arms, both kernel paths in one line), complete ncu launch list,
# DRAM traffic per launch, one full capture of each hot kernel, rows N1-N3, single-sequence latency, mid-size batches.
# Usage (repo root, under gpurun): bash tools/gpu_round.sh <tag>
T=${1:-r02}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu_$T.txt
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > $O/pytest_gpu_$T.txt; tail -2 $O/pytest_gpu_$T.txt
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 ) > $O/smoke_$T.txt; cat $O/smoke_$T.txt
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench_$T.json 2> $O/bench_$T.err; tail -c 300 $O/bench_$T.json; tail -3 $O/bench_$T.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_reference_$T.json 2>> $O/bench_$T.err
# every launch of one bench command (prepare ~80 launches, fp64 peak probe, warm-up, timed steps of both arms, e2e): -c 600 covers it
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$T.csv \
    python bench.py --steps 2 --warmup 3 --batch 592 --horizon 6 --no-cpu-baseline > $O/ncu_list_$T.log 2>&1
grep -c "uniform\|rollout_kernel" $O/launches_$T.csv
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:uniform_ -s 4 -c 2 --csv \
    --log-file $O/traffic_u_$T.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-general-path > /dev/null 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:rollout_kernel -s 2 -c 1 --csv \
    --log-file $O/traffic_g_$T.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --distinct-lengthscales > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:uniform_ -s 8 -c 2 -f -o $O/prof_$T \
    python bench.py --steps 1 --warmup 3 --batch 2368 --horizon 3 --no-cpu-baseline --no-general-path > $O/ncu_full_$T.log 2>&1
tail -1 $O/ncu_full_$T.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -f -o $O/prof_general_$T \
    python bench.py --steps 1 --warmup 3 --batch 1184 --horizon 3 --no-cpu-baseline --distinct-lengthscales > $O/ncu_full_general_$T.log 2>&1
tail -1 $O/ncu_full_general_$T.log | cut -c1-200
GPMPC_DEBUG_CLOCKS=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 2368 2>&1 >/dev/null | grep "gpmpc" | tail -4 > $O/phase_clocks_$T.txt; cut -c1-300 $O/phase_clocks_$T.txt
timeout 400 python tools/bench_next_rows.py 2>&1 | grep "^N[123]" > $O/next_rows_$T.txt; cat $O/next_rows_$T.txt
timeout 120 python tools/latency_single.py > $O/latency_$T.txt 2>&1; tail -6 $O/latency_$T.txt
timeout 200 python tools/bench_midbatch.py C4b > $O/midbatch_$T.txt 2>&1; tail -13 $O/midbatch_$T.txt
echo "# name | value | forward only | e2e | kernel ms | path" > $O/other_$T.txt
for w in C2 C3 C4a "C5 --batch 1184 --horizon 10"; do
  timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > $O/other_${T}_$(echo $w | cut -d' ' -f1).json 2>/dev/null
done
python tools/showbench.py $O/other_${T}_*.json >> $O/other_$T.txt; cat $O/other_$T.txt
python tools/showbench.py $O/bench_$T.json
