#!/bin/bash
# round-2 GPU call 2: static exp table + shared-memory column reduce (general + uniform kernels), issue-slot microbenchmarks
O=gpurun_out; T=r02b; mkdir -p $O
./tools/micro/issue_mix2 > $O/r02_micro_issue_mix2.txt 2>&1; cat $O/r02_micro_issue_mix2.txt
./tools/micro/issue_mix > $O/r02_micro_issue_mix.txt 2>&1; cat $O/r02_micro_issue_mix.txt
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/pytest_gpu_$T.txt; tail -3 $O/pytest_gpu_$T.txt
V=tools/micro/_variants
g() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --distinct-lengthscales --no-cpu-baseline --batch 2368 > $O/g_${T}_$name.json 2> $O/g_${T}_$name.err; }
u() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batch 2368 > $O/u_${T}_$name.json 2> $O/u_${T}_$name.err; }
g default X=1
g c2_192 GPMPC_GEN_CTAS=2
g clocks GPMPC_DEBUG_CLOCKS=1
grep "general clocks" $O/g_${T}_clocks.err | tail -2 | cut -c1-400
u default X=1
u bw192 GPMPC_LIB=$V/libgpmpc_bw192.so GPMPC_UNI_BWD_THREADS=192 GPMPC_UNI_BWD_CTAS=2
u bw192_nopremat GPMPC_LIB=$V/libgpmpc_bw192.so GPMPC_UNI_BWD_THREADS=192 GPMPC_UNI_BWD_CTAS=2 GPMPC_UNI_PREMAT=0
u clocks GPMPC_DEBUG_CLOCKS=1
grep "clocks/step" $O/u_${T}_clocks.err | tail -1 | cut -c1-400
python tools/showbench.py $O/g_${T}_*.json $O/u_${T}_*.json
tail -3 $O/g_${T}_default.err
