#!/bin/bash
# round-2 GPU call 5 (2 GPUs): NCCL parity test, new bench.py line (general_path, roofline, checksum) at N=1 and N=2, table addressing A/B
O=gpurun_out; T=r02e; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$T.txt 2>&1; grep -v "^frame\|^#" $O/pytest_gpu_$T.txt | tail -6 | cut -c1-300
V=tools/micro/_variants
timeout 600 python bench.py --steps 2 --warmup 3 --batch 2368 --no-cpu-baseline > $O/bench_$T.json 2> $O/bench_$T.err; tail -3 $O/bench_$T.err
GPMPC_LIB=$V/libgpmpc_immoff0.so timeout 600 python bench.py --steps 2 --warmup 3 --batch 2368 --no-cpu-baseline > $O/bench_${T}_immoff0.json 2> $O/bench_${T}_immoff0.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --batch 2368 --no-cpu-baseline > $O/bench_${T}_2gpu.json 2> $O/bench_${T}_2gpu.err; tail -3 $O/bench_${T}_2gpu.err
python tools/showbench.py $O/bench_$T.json $O/bench_${T}_immoff0.json $O/bench_${T}_2gpu.json
python - <<'PY'
import json
for f in ("gpurun_out/bench_r02e.json", "gpurun_out/bench_r02e_2gpu.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, "checksum", d["cost_checksum"], "argmin", d["cost_argmin"], "frac", round(r["frac"], 3), [round(k["frac"], 3) for k in r["kernels"]],
              "general", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.get("general_path", {}).items() if k in ("value", "forward_only", "executed_frac", "ceiling_predictions_per_s", "cost_checksum")})
    except Exception as e:
        print(f, "unreadable", e)
PY
