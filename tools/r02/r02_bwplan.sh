#!/bin/bash
# reverse sweep: 3 x 128 threads (default when it fits) vs 2 x 256 by batch size
O=gpurun_out; T=${1:-r03j}; mkdir -p $O
for w in "C4b --batch 1024" "C4b --batch 2048" "C4b --batch 4096" "C2" "C3" "C4a"; do
  tag=$(echo $w | tr -d ' -' )
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-general-path > $O/p_${T}_${tag}_3x128.json 2>/dev/null
  GPMPC_UNI_BWD_CTAS=2 GPMPC_UNI_BWD_THREADS=256 GPMPC_UNI_PREMAT=1 timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-general-path > $O/p_${T}_${tag}_2x256.json 2>/dev/null
done
python tools/showbench.py $O/p_${T}_*.json
