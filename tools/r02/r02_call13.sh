#!/bin/bash
# round-2 GPU call 13: iK loads without L1 allocation, unroll-2 record building
O=gpurun_out; T=r02m; mkdir -p $O
V=tools/micro/_variants
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$T.txt 2>&1; grep -v "^frame\|^#" $O/pytest_gpu_$T.txt | tail -3 | cut -c1-300
u() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --batch 2368 > $O/u_${T}_$name.json 2> $O/u_${T}_$name.err; }
u default X=1
u ikl1 GPMPC_LIB=$V/libgpmpc_ikl1.so
u clocks GPMPC_DEBUG_CLOCKS=1
grep "clocks/step" $O/u_${T}_clocks.err | tail -1 | cut -c1-400
u default2 X=1
u ikl1_2 GPMPC_LIB=$V/libgpmpc_ikl1.so
python tools/showbench.py $O/u_${T}_*.json
