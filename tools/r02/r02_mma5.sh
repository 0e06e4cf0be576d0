#!/bin/bash
# x loads of the record phases one stage ahead (UNI_X_PIPELINE) + diagonal-pair beta in the column record (general kernel)
O=gpurun_out; T=${1:-r02y}; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > $O/pytest_gpu_$T.txt; tail -1 $O/pytest_gpu_$T.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_$T.json 2> $O/bench_$T.err
python tools/showbench.py $O/bench_$T.json
GPMPC_DEBUG_CLOCKS=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-general-path --batch 2368 2>&1 >/dev/null | grep "gpmpc clocks" | tail -1 | cut -c1-300
