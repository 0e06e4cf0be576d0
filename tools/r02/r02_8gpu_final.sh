#!/bin/bash
# final 8-GPU runs of round 2 with the final binaries: C4b sharded over 8 GPUs, BASELINE.json config 5 for real
O=gpurun_out; T=${1:-r02w8}; mkdir -p $O
run() { n=$1; tag=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $n "$@" > $O/bench_${T}_$tag.json 2> $O/bench_${T}_$tag.err; tail -2 $O/bench_${T}_$tag.err | cut -c1-200; }
run 8 c4b_8gpu --steps 3 --warmup 3 --no-cpu-baseline
run 8 c5_8gpu --workload C5 --steps 2 --warmup 3 --no-cpu-baseline --no-general-path
python tools/showbench.py $O/bench_${T}_*.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${T}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "n_gpus", d["n_gpus"], "value %.0f e2e %.0f ms %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), "checksum", d["cost_checksum"], "argmin", d["cost_argmin"],
              "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "general", round((d.get("general_path") or {}).get("value", 0)), "frac", [round(k["frac"], 3) for k in d["roofline"]["kernels"]])
    except Exception as e:
        print(f, "unreadable", e)
PY
