#!/bin/bash
# round-2 GPU call 14: general kernel, off-diagonal pairs with four rows per lane (gen_cols4)
O=gpurun_out; T=r02n; mkdir -p $O
V=tools/micro/_variants
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$T.txt 2>&1; grep -v "^frame\|^#" $O/pytest_gpu_$T.txt | tail -3 | cut -c1-300
g() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --distinct-lengthscales --no-cpu-baseline --batch 2368 > $O/g_${T}_$name.json 2> $O/g_${T}_$name.err; }
g rows4 X=1
g rows2 GPMPC_LIB=$V/libgpmpc_rows2.so
g rows4_clocks GPMPC_DEBUG_CLOCKS=1
grep "general" $O/g_${T}_rows4_clocks.err | tail -2 | cut -c1-400
g rows4_512 GPMPC_GEN_THREADS=384
timeout 300 python bench.py --workload C2 --steps 2 --warmup 3 --no-cpu-baseline --distinct-lengthscales > $O/g_${T}_c2.json 2> $O/g_${T}_c2.err
timeout 300 python bench.py --workload C5 --steps 2 --warmup 3 --no-cpu-baseline --batch 592 --horizon 10 --distinct-lengthscales > $O/g_${T}_c5.json 2> $O/g_${T}_c5.err; tail -2 $O/g_${T}_c5.err | cut -c1-200
python tools/showbench.py $O/g_${T}_*.json
