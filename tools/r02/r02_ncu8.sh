#!/bin/bash
# source-level ncu capture of the E = 8 uniform kernels (C5 shard)
O=gpurun_out; T=${1:-r03b}; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:uniform_ -s 4 -c 2 -f -o $O/prof8_$T \
    python bench.py --workload C5 --steps 1 --warmup 2 --batch 592 --horizon 4 --no-cpu-baseline --no-general-path > $O/ncu8_$T.log 2>&1
tail -2 $O/ncu8_$T.log | cut -c1-200
