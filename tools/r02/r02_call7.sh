#!/bin/bash
# round-2 GPU call 7: warp-cooperative B0/B4 of the uniform reverse sweep
O=gpurun_out; T=r02g; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$T.txt 2>&1; grep -v "^frame\|^#" $O/pytest_gpu_$T.txt | tail -4 | cut -c1-300
u() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-general-path --batch 2368 > $O/u_${T}_$name.json 2> $O/u_${T}_$name.err; }
u default X=1
u clocks GPMPC_DEBUG_CLOCKS=1
grep "clocks/step" $O/u_${T}_clocks.err | tail -1 | cut -c1-400
u nopremat GPMPC_UNI_PREMAT=0
timeout 300 python bench.py --workload C5 --steps 2 --warmup 3 --no-cpu-baseline --no-general-path --batch 1184 --horizon 10 > $O/u_${T}_c5.json 2> $O/u_${T}_c5.err
timeout 300 python bench.py --workload C2 --steps 2 --warmup 3 --no-cpu-baseline --no-general-path > $O/u_${T}_c2.json 2> $O/u_${T}_c2.err
timeout 300 python bench.py --workload C3 --steps 2 --warmup 3 --no-cpu-baseline --no-general-path > $O/u_${T}_c3.json 2> $O/u_${T}_c3.err
timeout 300 python bench.py --workload C4a --steps 2 --warmup 3 --no-cpu-baseline --no-general-path > $O/u_${T}_c4a.json 2> $O/u_${T}_c4a.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_$T.json 2> $O/bench_$T.err
python tools/showbench.py $O/u_${T}_*.json $O/bench_$T.json
