#!/bin/bash
# full GPU test-suite, smoke, headline bench, the other BASELINE.json shapes
O=gpurun_out; T=${1:-r03i}; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > $O/pytest_gpu_$T.txt; tail -1 $O/pytest_gpu_$T.txt
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 ) > $O/smoke_$T.txt; cat $O/smoke_$T.txt
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_$T.json 2> $O/bench_$T.err
echo "# name | value | forward only | e2e | kernel ms | path" > $O/other_$T.txt
for w in C1 C2 C3 C4a "C5 --batch 1184 --horizon 10"; do
  timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > $O/other_${T}_$(echo $w | cut -d' ' -f1).json 2>/dev/null
done
python tools/showbench.py $O/other_${T}_*.json >> $O/other_$T.txt; cat $O/other_$T.txt
python tools/showbench.py $O/bench_$T.json
timeout 120 python tools/latency_single.py > $O/latency_$T.txt 2>&1; tail -5 $O/latency_$T.txt
timeout 200 python tools/bench_midbatch.py C4b > $O/midbatch_$T.txt 2>&1; tail -12 $O/midbatch_$T.txt
