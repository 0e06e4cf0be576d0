#!/bin/bash
# round-2 GPU call 12: tail split (1 GPU, B = 1024 = one rank's shard at 8 GPUs)
O=gpurun_out; T=r02l; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$T.txt 2>&1; grep -v "^frame\|^#" $O/pytest_gpu_$T.txt | tail -4 | cut -c1-300
u() { name=$1; b=$2; shift 2; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-general-path --batch $b > $O/u_${T}_$name.json 2> $O/u_${T}_$name.err; }
u b1024_split 1024 X=1
u b1024_nosplit 1024 GPMPC_UNI_NO_TAIL_SPLIT=1
u b2048_split 2048 X=1
u b2048_nosplit 2048 GPMPC_UNI_NO_TAIL_SPLIT=1
u b4096_split 4096 X=1
u b4096_nosplit 4096 GPMPC_UNI_NO_TAIL_SPLIT=1
u b600_split 600 X=1
u b600_nosplit 600 GPMPC_UNI_NO_TAIL_SPLIT=1
u b8192 8192 X=1
python tools/showbench.py $O/u_${T}_*.json
