#!/bin/bash
# ncu artefacts of the FINAL library: launch list of one bench command, DRAM traffic and a full capture of the uniform kernels
O=gpurun_out; T=${1:-r03n}; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$T.csv \
    python bench.py --steps 2 --warmup 3 --batch 592 --horizon 6 --no-cpu-baseline > $O/ncu_list_$T.log 2>&1
grep -c "uniform\|rollout_kernel" $O/launches_$T.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:uniform_ -c 40 --csv --log-file $O/launches_big_$T.csv \
    python bench.py --steps 2 --warmup 3 --batch 2368 --horizon 6 --no-cpu-baseline --no-general-path > $O/ncu_list_big_$T.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:uniform_ -s 4 -c 2 --csv \
    --log-file $O/traffic_u_$T.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-general-path > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:uniform_ -s 8 -c 2 -f -o $O/prof_$T \
    python bench.py --steps 1 --warmup 3 --batch 2368 --horizon 3 --no-cpu-baseline --no-general-path > $O/ncu_full_$T.log 2>&1
tail -1 $O/ncu_full_$T.log | cut -c1-200
