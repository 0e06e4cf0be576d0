#!/bin/bash
# Follow-up of gpu_round_s8.sh: complete ncu launch list of one bench command (prepare + warm-up + timed steps + e2e),
# the other BASELINE.json workloads, and the rows around the hot path.
TAG=${1:-s8}
O=gpurun_out
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --batch 592 --horizon 6 --no-cpu-baseline > $O/ncu_list_$TAG.log 2>&1
grep -c uniform $O/launches_$TAG.csv
echo "# other BASELINE.json workloads on one B200 (bench.py --workload <name> --steps 2 --warmup 3 --no-cpu-baseline; C5: one 1/8 shard-sized batch, --batch 1184 --horizon 10)" > $O/other_$TAG.txt
echo "# name | workload | predictions/s objective+gradient | objective only | kernel ms | prepare ms | fp64 executed frac fwd/bwd" >> $O/other_$TAG.txt
for w in C2 C3 C4a "C5 --batch 1184 --horizon 10"; do
  timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); f=d['roofline']['fp64']
print('$w'.split()[0], '|', d['config']['workload'], '|', round(d['value']), '|', round(d['forward_only']['value']), '|', d['kernel_ms'], '|', d['prepare_ms']['steady'], '| %.2f/%.2f' % (f['executed_frac_fwd'], f.get('executed_frac_bwd', 0)))" >> $O/other_$TAG.txt
done
cat $O/other_$TAG.txt
timeout 300 python tools/bench_next_rows.py > $O/next_rows_$TAG.txt 2>&1; tail -9 $O/next_rows_$TAG.txt
timeout 120 python tools/latency_single.py > $O/latency_$TAG.txt 2>&1; tail -8 $O/latency_$TAG.txt
