#!/bin/bash
# warp-cooperative E x E inverses in the forward P0: parity, headline, C5 shard, phase clocks
O=gpurun_out; T=${1:-r03c}; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > $O/pytest_gpu_$T.txt; tail -1 $O/pytest_gpu_$T.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_$T.json 2> $O/bench_$T.err
timeout 300 python bench.py --workload C5 --batch 1184 --horizon 10 --steps 2 --warmup 3 --no-cpu-baseline --no-general-path > $O/c5_$T.json 2> $O/c5_$T.err
python tools/showbench.py $O/bench_$T.json $O/c5_$T.json
GPMPC_DEBUG_CLOCKS=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-general-path --batch 2368 2>&1 >/dev/null | grep "gpmpc clocks" | tail -1 | cut -c1-300
GPMPC_DEBUG_CLOCKS=1 timeout 300 python bench.py --workload C5 --batch 592 --horizon 10 --steps 1 --warmup 3 --no-cpu-baseline --no-general-path 2>&1 >/dev/null | grep "gpmpc clocks" | tail -1 | cut -c1-300
