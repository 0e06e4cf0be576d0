#!/bin/bash
# Final GPU-box round of this session (trimmed gpu_round.sh): parity tests, smoke, bench (both arms, both kernel paths),
# launch-plan variants, ncu launch list, DRAM traffic per launch, one full capture of the uniform kernels.
TAG=${1:-s8}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu_$TAG.txt
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/pytest_gpu_$TAG.txt
tail -1 $O/pytest_gpu_$TAG.txt
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 ) > $O/smoke_$TAG.txt
cat $O/smoke_$TAG.txt
timeout 600 python bench.py --steps 3 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err
tail -c 300 $O/bench_$TAG.json; tail -3 $O/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_reference_$TAG.json 2>> $O/bench_$TAG.err
timeout 600 python bench.py --steps 2 --warmup 3 --distinct-lengthscales --no-cpu-baseline > $O/bench_distinct_$TAG.json 2>> $O/bench_$TAG.err
for v in "3 128"; do set -- $v
  GPMPC_UNI_FWD_CTAS=$1 GPMPC_UNI_FWD_THREADS=$2 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fwd plan $1 x $2:', d['kernel_ms'], d['value'])" | tee -a $O/plans_$TAG.txt
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --batch 592 --horizon 6 --no-cpu-baseline > $O/ncu_list_$TAG.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:uniform_ -s 4 -c 2 --csv \
    --log-file $O/traffic_u_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:uniform_ -s 8 -c 2 -f -o $O/prof_$TAG \
    python bench.py --steps 1 --warmup 3 --batch 2368 --horizon 3 --no-cpu-baseline > $O/ncu_full_$TAG.log 2>&1
tail -1 $O/ncu_full_$TAG.log | cut -c1-200
ls -la $O | tail -12
