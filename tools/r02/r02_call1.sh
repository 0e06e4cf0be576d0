#!/bin/bash
# round-2 GPU call 1: parity of the rebuilt general-path kernel + launch-plan variants; uniform reverse sweep 3 x 128 variant
O=gpurun_out; T=r02a; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/pytest_gpu_$T.txt; tail -3 $O/pytest_gpu_$T.txt
V=tools/micro/_variants
g() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --distinct-lengthscales --no-cpu-baseline --batch 2368 > $O/g_${T}_$name.json 2> $O/g_${T}_$name.err; }
u() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batch 2368 > $O/u_${T}_$name.json 2> $O/u_${T}_$name.err; }
g default X=1
g c1_384 GPMPC_GEN_CTAS=1
g c2_128 GPMPC_GEN_CTAS=2 GPMPC_GEN_THREADS=128
g c2_160 GPMPC_GEN_CTAS=2 GPMPC_GEN_THREADS=160
g seg16 GPMPC_GEN_SEG=16
g seg64 GPMPC_GEN_SEG=64
g v512_c2_256 GPMPC_LIB=$V/libgpmpc_g512.so
g v512_c1_512 GPMPC_LIB=$V/libgpmpc_g512.so GPMPC_GEN_CTAS=1
g clocks GPMPC_DEBUG_CLOCKS=1
grep "general clocks" $O/g_${T}_clocks.err | tail -2 | cut -c1-400
u default X=1
u bw3 GPMPC_LIB=$V/libgpmpc_bw3.so GPMPC_UNI_PREMAT=0 GPMPC_UNI_BWD_THREADS=128 GPMPC_UNI_BWD_CTAS=3
u bw3_premat GPMPC_LIB=$V/libgpmpc_bw3.so GPMPC_UNI_BWD_THREADS=128 GPMPC_UNI_BWD_CTAS=3
python tools/showbench.py $O/g_${T}_*.json $O/u_${T}_*.json
tail -3 $O/g_${T}_default.err
