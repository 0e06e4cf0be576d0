#!/bin/bash
# DMMA forward sweep (E = 4): parity of the uniform path + forward kernel time at the headline shape
O=gpurun_out; T=${1:-r02s}; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6 ) > $O/pytest_gpu_$T.txt; tail -3 $O/pytest_gpu_$T.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-general-path > $O/bench_$T.json 2> $O/bench_$T.err; tail -2 $O/bench_$T.err
python tools/showbench.py $O/bench_$T.json
GPMPC_DEBUG_CLOCKS=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-general-path --batch 2368 2>&1 >/dev/null | grep "gpmpc" | tail -4 | cut -c1-400
