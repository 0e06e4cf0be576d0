#!/bin/bash
# round-2 GPU call 15: general kernel variants (rows per lane x iK load policy), C5 full-size gradient on the general path
O=gpurun_out; T=r02o; mkdir -p $O
V=tools/micro/_variants
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$T.txt 2>&1; grep -v "^frame\|^#" $O/pytest_gpu_$T.txt | tail -3 | cut -c1-300
g() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --distinct-lengthscales --no-cpu-baseline --batch 2368 > $O/g_${T}_$name.json 2> $O/g_${T}_$name.err; }
g rows4_pf X=1
g rows2_pf GPMPC_LIB=$V/libgpmpc_rows2pf.so
g rows4_nopf GPMPC_LIB=$V/libgpmpc_rows4nopf.so
g rows4_pf_b X=1
g rows2_pf_b GPMPC_LIB=$V/libgpmpc_rows2pf.so
timeout 300 python bench.py --workload C5 --steps 2 --warmup 3 --no-cpu-baseline --batch 296 --horizon 10 --distinct-lengthscales > $O/g_${T}_c5.json 2> $O/g_${T}_c5.err; tail -2 $O/g_${T}_c5.err | cut -c1-200
python tools/showbench.py $O/g_${T}_*.json
