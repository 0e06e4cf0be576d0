#!/bin/bash
# round-2 GPU call 10: fused L-BFGS update kernel
O=gpurun_out; T=r02i; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$T.txt 2>&1; grep -v "^frame\|^#" $O/pytest_gpu_$T.txt | tail -6 | cut -c1-300
timeout 600 python tools/bench_next_rows.py 2>&1 | grep "^N1" > $O/next_rows_$T.txt; cat $O/next_rows_$T.txt
