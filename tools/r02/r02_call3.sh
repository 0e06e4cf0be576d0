#!/bin/bash
# round-2 GPU call 3: full pytest log; sign-through-magic-constant general kernel; uniform reverse sweep: smem vs shuffle column reduce
O=gpurun_out; T=r02c; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$T.txt 2>&1; grep -v "^frame\|^#" $O/pytest_gpu_$T.txt | tail -25 | cut -c1-300
V=tools/micro/_variants
g() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --distinct-lengthscales --no-cpu-baseline --batch 2368 > $O/g_${T}_$name.json 2> $O/g_${T}_$name.err; }
u() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batch 2368 > $O/u_${T}_$name.json 2> $O/u_${T}_$name.err; }
g default X=1
g clocks GPMPC_DEBUG_CLOCKS=1
grep "general" $O/g_${T}_clocks.err | tail -3 | cut -c1-400
u default X=1
u bwshfl GPMPC_LIB=$V/libgpmpc_bwshfl.so
u bwshfl_t128 GPMPC_LIB=$V/libgpmpc_bwshfl.so GPMPC_UNI_BWD_THREADS=128 GPMPC_UNI_BWD_CTAS=2
u smem_c1 GPMPC_UNI_BWD_CTAS=1
python tools/showbench.py $O/g_${T}_*.json $O/u_${T}_*.json
tail -3 $O/g_${T}_default.err
