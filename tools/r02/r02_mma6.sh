#!/bin/bash
# product (x pipeline off) vs xpipe variant (volatile early loads)
O=gpurun_out; T=${1:-r02z}; mkdir -p $O
V=tools/micro/_variants
u() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-general-path > $O/u_${T}_$name.json 2> $O/u_${T}_$name.err; }
u default X=1
u xpipe GPMPC_LIB=$V/libgpmpc_xpipe.so
python tools/showbench.py $O/u_${T}_*.json
GPMPC_LIB=$V/libgpmpc_xpipe.so GPMPC_DEBUG_CLOCKS=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-general-path --batch 2368 2>&1 >/dev/null | grep "gpmpc clocks" | tail -1 | cut -c1-300
( GPMPC_LIB=$V/libgpmpc_xpipe.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 ) > $O/pytest_gpu_$T.txt; tail -1 $O/pytest_gpu_$T.txt
