#!/bin/bash
# round-2 GPU call 8: general-path phase trims (P2 unroll, column coefficients), mid-size batch clusters
O=gpurun_out; T=r02h; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$T.txt 2>&1; grep -v "^frame\|^#" $O/pytest_gpu_$T.txt | tail -4 | cut -c1-300
g() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --distinct-lengthscales --no-cpu-baseline --batch 2368 > $O/g_${T}_$name.json 2> $O/g_${T}_$name.err; }
g default X=1
g clocks GPMPC_DEBUG_CLOCKS=1
grep "general clocks" $O/g_${T}_clocks.err | tail -1 | cut -c1-400
timeout 300 python tools/bench_midbatch.py C4b > $O/midbatch_${T}_cap2.txt 2>&1; cat $O/midbatch_${T}_cap2.txt
GPMPC_UNI_CLUSTER_CAP=1 timeout 300 python tools/bench_midbatch.py C4b > $O/midbatch_${T}_cap1.txt 2>&1; cat $O/midbatch_${T}_cap1.txt
timeout 300 python tools/bench_midbatch.py C4b --distinct > $O/midbatch_${T}_distinct.txt 2>&1; cat $O/midbatch_${T}_distinct.txt
timeout 400 python tools/bench_next_rows.py 2>&1 | grep "^N[12]" > $O/next_rows_$T.txt; cat $O/next_rows_$T.txt
python tools/showbench.py $O/g_${T}_*.json
