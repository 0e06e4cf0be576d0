"""Tuning sweep for the uniform-kernel launch plan (threads / CTAs per SM / item segment sizes)."""
import os
import subprocess
import sys

CODE = r'''
import sys, os, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "data-efficient-reinforcement-learning-with-probabilistic-model-predictive-control_b200"))
from tools.gpu_check import engine_for, run_rollout
from oracle.workloads import make_workload
cfg = make_workload("C4b", B=2368, H=6)
eng = engine_for(cfg); eng.enable_timing(True)
for g in (False, True):
    for it in range(2):
        run_rollout(eng, cfg, need_grad=g); torch.cuda.synchronize()
    f, b = eng.last_rollout_ms(), max(eng.last_backward_ms(), 0.0)
    print("grad=%d fwd %.2f bwd %.2f ms -> %.0f preds/s" % (g, f, b, cfg["B"] * cfg["H"] / (f + b) * 1e3), end=" | ")
print()
'''
grid = [{}, {"GPMPC_UNI_PREMAT": 0}]
for cfg in grid:
    env = dict(os.environ)
    env.update({k: str(v) for k, v in cfg.items()})
    out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print(cfg, "->", out.stdout.strip()[-200:], out.stderr.strip()[-300:] if out.returncode else "", flush=True)
