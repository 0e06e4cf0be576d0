#!/bin/bash
# tensor-core sweeps for E >= 6 (UNI_MMA8): parity + the C5 shard
O=gpurun_out; T=${1:-r02v}; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6 ) > $O/pytest_gpu_$T.txt; tail -3 $O/pytest_gpu_$T.txt
timeout 300 python bench.py --workload C5 --batch 1184 --horizon 10 --steps 2 --warmup 3 --no-cpu-baseline --no-general-path > $O/c5_$T.json 2> $O/c5_$T.err; tail -2 $O/c5_$T.err
python tools/showbench.py $O/c5_$T.json
GPMPC_DEBUG_CLOCKS=1 timeout 300 python bench.py --workload C5 --batch 592 --horizon 10 --steps 1 --warmup 3 --no-cpu-baseline --no-general-path 2>&1 >/dev/null | grep "gpmpc clocks" | tail -1 | cut -c1-300
