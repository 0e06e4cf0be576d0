#!/bin/bash
# round-2 GPU call 16: which of (iK loads without L1 allocation, two points in flight in P1a) regressed C3 (E=2) and the C5 shard (E=8)?
O=gpurun_out; T=r02q; mkdir -p $O
V=tools/micro/_variants
w() { name=$1; lib=$2; shift 2; GPMPC_LIB=$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-general-path "$@" > $O/u_${T}_$name.json 2> $O/u_${T}_$name.err; }
D=data-efficient-reinforcement-learning-with-probabilistic-model-predictive-control_b200/rl_gp_mpc/_lib/libgpmpc.so
for v in default ikl1 nounroll; do
  lib=$D; [ $v != default ] && lib=$V/libgpmpc_$v.so
  w ${v}_c3 $lib --workload C3
  w ${v}_c5 $lib --workload C5 --batch 1184 --horizon 10
  w ${v}_c4b $lib --batch 2368
  w ${v}_c2 $lib --workload C2
done
python tools/showbench.py $O/u_${T}_*.json
