#!/bin/bash
# reverse sweep as 3 CTAs x 128 threads (168 registers) with the per-step records in the global scratch (premat = 2)
O=gpurun_out; T=${1:-r03h}; mkdir -p $O
V=tools/micro/_variants
u() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-general-path > $O/u_${T}_$name.json 2> $O/u_${T}_$name.err; tail -1 $O/u_${T}_$name.err | cut -c1-150; }
u default X=1
u premat2 GPMPC_UNI_PREMAT=2
u bw3 GPMPC_LIB=$V/libgpmpc_bw3.so GPMPC_UNI_PREMAT=2 GPMPC_UNI_BWD_THREADS=128 GPMPC_UNI_BWD_CTAS=3
u bw3_2x GPMPC_LIB=$V/libgpmpc_bw3.so GPMPC_UNI_PREMAT=2 GPMPC_UNI_BWD_THREADS=128 GPMPC_UNI_BWD_CTAS=2
python tools/showbench.py $O/u_${T}_*.json
