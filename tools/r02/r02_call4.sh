#!/bin/bash
# round-2 GPU call 4: balanced two-list split + iK prefetch (general), uniform reverse sweep back on shuffles
O=gpurun_out; T=r02d; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$T.txt 2>&1; grep -v "^frame\|^#" $O/pytest_gpu_$T.txt | tail -12 | cut -c1-300
V=tools/micro/_variants
g() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --distinct-lengthscales --no-cpu-baseline --batch 2368 > $O/g_${T}_$name.json 2> $O/g_${T}_$name.err; }
u() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batch 2368 > $O/u_${T}_$name.json 2> $O/u_${T}_$name.err; }
g default X=1
g nopf GPMPC_LIB=$V/libgpmpc_nopf.so
g clocks GPMPC_DEBUG_CLOCKS=1
grep "general" $O/g_${T}_clocks.err | tail -2 | cut -c1-400
g seg16 GPMPC_GEN_SEG=16
g seg64 GPMPC_GEN_SEG=64
u default X=1
timeout 300 python bench.py --workload C5 --steps 2 --warmup 3 --no-cpu-baseline --batch 1184 --horizon 10 > $O/u_${T}_c5.json 2> $O/u_${T}_c5.err
timeout 300 python bench.py --workload C5 --steps 2 --warmup 3 --no-cpu-baseline --batch 592 --horizon 10 --distinct-lengthscales > $O/g_${T}_c5.json 2> $O/g_${T}_c5.err
timeout 300 python bench.py --workload C2 --steps 2 --warmup 3 --no-cpu-baseline --distinct-lengthscales > $O/g_${T}_c2.json 2> $O/g_${T}_c2.err
python tools/showbench.py $O/g_${T}_*.json $O/u_${T}_*.json
tail -3 $O/g_${T}_default.err
