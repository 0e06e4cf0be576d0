#!/bin/bash
# round-2 GPU call 11 (8 GPUs): BASELINE.json config 5 for real (E=8, N=1000, H=50, B=65536 over 8 GPUs), C4b at 8 / 4 / 1 GPUs
O=gpurun_out; T=r02k; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $O/gpus_$T.txt; head -3 $O/gpus_$T.txt
run() { n=$1; tag=$2; shift 2; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $n "$@" > $O/bench_${T}_$tag.json 2> $O/bench_${T}_$tag.err; tail -2 $O/bench_${T}_$tag.err | cut -c1-200; }
run 8 c4b_8gpu --steps 3 --warmup 3 --no-cpu-baseline
NCCL_DEBUG=INFO run 8 c5_8gpu --workload C5 --steps 2 --warmup 3 --no-cpu-baseline --no-general-path
grep -i "nvls\|NCCL version\|via P2P\|nChannels" $O/bench_${T}_c5_8gpu.err | sort | uniq -c | sort -rn | head -8 > $O/nccl_info_$T.txt; cat $O/nccl_info_$T.txt | cut -c1-200
run 4 c4b_4gpu --steps 3 --warmup 3 --no-cpu-baseline
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_${T}_c4b_1gpu.json 2> $O/bench_${T}_c4b_1gpu.err
python tools/showbench.py $O/bench_${T}_*.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r02k_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "n_gpus", d["n_gpus"], "value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]), "checksum", d["cost_checksum"], "argmin", d["cost_argmin"],
              "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "general", round(d.get("general_path", {}).get("value", 0)))
    except Exception as e:
        print(f, "unreadable", e)
PY
