#!/usr/bin/env python
"""bench.py -- horizon-step GP predictions/sec of the GP-MPC inner loop (BASELINE.json metric).

    python bench.py --gpus 1 --steps 3 --warmup 3                       # this framework (CUDA, sm_100a)
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 3      # the reference algorithm on host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N

A "step" is ONE batched evaluation of the MPC objective AND its gradient (the reference's
compute_mean_lcb_trajectory, controllers/gp_mpc_controller.py:229-285) for the whole candidate batch of the
workload BASELINE.json quotes the metric on: ProcessControl-shaped "4-state/2-action" GP (E=4, Na=2, D=6),
N=500 stored transitions, horizon H=30, B=8192 candidate action sequences (oracle/workloads.py "C4b").
predictions = B*H per step (one prediction = one predict_next_state_change, gp_model.py:112).
With N GPUs the SAME batch of 8192 candidates is sharded over the ranks (strong scaling) and the per-candidate
costs are all-gathered over NCCL inside the timed region.

The JSON line carries: value (device-resident inputs, fwd+grad), forward_only (same without the gradient),
e2e (host buffers through GpMpcController.compute_mean_lcb_trajectory_batch, H2D + D2H inside the timed
region), roofline (algorithmic bytes per SURVEY.md 8(d) over the rollout kernel's CUDA-event time, against
MEASURED_PEAKS.json; plus the float64-FMA fraction, which is the real bound), cpu_baseline (oracle port on
host cores, bounded sample), clocks (nvidia-smi sampled during the timed region), gpu_launches.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "data-efficient-reinforcement-learning-with-probabilistic-model-predictive-control_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle.workloads import (algorithmic_bytes_per_prediction, algorithmic_flops_per_prediction,  # noqa: E402
                              full_lengthscale, make_workload)

METRIC = "horizon-step GP predictions/sec at N=500, H=30, batch=8192"
UNIT = "predictions/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C4b")
    ap.add_argument("--batch", type=int, default=None, help="override the candidate batch (debug only)")
    ap.add_argument("--horizon", type=int, default=None, help="override the horizon (debug only)")
    ap.add_argument("--distinct-lengthscales", action="store_true",
                    help="per-GP ARD lengthscales ('trained' hyper-parameters) instead of the reference defaults")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


_STDOUT_FD = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _STDOUT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_STDOUT_FD, data)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_port_rate(cfg, candidates, horizon, need_grad=True, repeats=1):
    """Oracle (float64 torch CPU port of the reference algorithm) on a bounded sample; predictions/s."""
    from oracle import gpmpc_oracle as orc
    sub = dict(cfg)
    sub["H"] = horizon
    sub["actions"] = cfg["actions"][:, :horizon]
    model = orc.model_from_workload(sub)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        orc.evaluate_workload(sub, candidates=list(range(candidates)), need_grad=need_grad, model=model)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return candidates * horizon / best, best


def run_reference(args, cfg, rank):
    """--impl reference: the reference's algorithm (oracle port; the Python reference cannot travel to the
    GPU box) on the host cores, same metric/unit/config.  Each step = a bounded sample of the workload."""
    if rank != 0:
        return
    try:                                  # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses every host core it may
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    cand, hor = min(4, cfg["B"]), cfg["H"]      # ~3 s of host work per step at the headline shape
    for _ in range(args.warmup):
        cpu_port_rate(cfg, 1, min(2, hor))
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):
        rate, dt = cpu_port_rate(cfg, cand, hor)
        n += cand * hor
    total = time.perf_counter() - t0
    value = n / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg, args),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d candidate x %d horizon steps per step, objective+autograd gradient "
                                       "(oracle/gpmpc_oracle.py), float64" % (cand, hor)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(cfg, args):
    return {"workload": "%s: E=%d Na=%d D=%d N=%d H=%d B=%d (candidate batch sharded over ranks)" % (
        cfg["name"], cfg["E"], cfg["Na"], cfg["D"], cfg["N"], cfg["H"], cfg["B"]),
        "hyperparameters": "distinct per-GP ARD lengthscales" if args.distinct_lengthscales
        else "reference defaults (lengthscale %.2f, outputscale 5e-2, noise 1e-5)" % cfg["lengthscale"][0, 0],
        "step": "one objective+gradient evaluation of the whole batch (compute_mean_lcb_trajectory)",
        "l2": "inputs >> L2 not applicable: the 8 MB training block is L2-resident by design; per-step records "
              "(1.2 GB per step) exceed L2, and a 256 MB buffer is rewritten between timed iterations (L2 flush)"}


def fp64_report(uniform, E, N, preds, fwd_ms, bwd_ms, f_alg, peak):
    """Float64 roofline: (a) SURVEY 8(d)'s algorithmic flops (E^2-pair algorithm) and (b) the float64 instructions the
    kernels actually execute per prediction in their hot loops (counted in the SASS), each counted
    as one FMA = 2 flops, against the measured DFMA peak.  NB: a DFMA with three distinct register operands issues at
    2/3 of that peak on B200 (tools/micro/dfma_operands.cu)."""
    NP = (N + 63) // 64 * 64
    if uniform:
        tri = (NP // 64) * (NP // 64 + 1) // 2 * 64 * 64      # elements of the upper tile triangle (both sweeps)
        # (the row factor of the exponential is applied after the sweep: E exponent FMAs, no per-element add)
        fwd_ops = tri * (E + 7 + E + 1)                       # exponent, exp2s, beta-weighted row sums, trace
        bwd_ops = tri * (E + 7 + (E + 1) + 3.06 + E)          # + coefficient, w, rho/col sums, xi
    else:
        elems = E * NP * (NP + 64) // 2 + E * (E - 1) // 2 * NP * NP
        fwd_ops = elems * ((E + 1) + 7 + 2 + 1 + (E + 3.1))   # gradient mode: + rho/gamma/xi accumulation
        bwd_ops = 0
    ex_f = 2.0 * fwd_ops * preds / (fwd_ms * 1e-3)
    out = {"peak_tflops": peak / 1e12, "peak_source": "measured (gpmpc_fp64_peak: register-resident DFMA loop)",
           "algorithmic_flops_per_prediction": f_alg,
           "algorithmic_tflops": preds * f_alg / (fwd_ms * 1e-3) / 1e12,
           "algorithmic_frac": preds * f_alg / (fwd_ms * 1e-3) / peak,
           "executed_flops_per_prediction_fwd": 2.0 * fwd_ops, "executed_tflops_fwd": ex_f / 1e12,
           "executed_frac_fwd": ex_f / peak}
    if bwd_ops and bwd_ms > 0:
        ex_b = 2.0 * bwd_ops * preds / (bwd_ms * 1e-3)
        out.update({"executed_flops_per_prediction_bwd": 2.0 * bwd_ops, "executed_tflops_bwd": ex_b / 1e12,
                    "executed_frac_bwd": ex_b / peak})
    return out


def ncu_traffic(kernel_prefix, workload):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this workload
    (profiles/dram_traffic.json, written by tools/ncu_traffic.py from `ncu --metrics dram__bytes_read.sum,
    dram__bytes_write.sum`); None when no capture of this kernel/workload is committed."""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
        for row in tab["launches"]:
            if row["kernel"].startswith(kernel_prefix) and row["workload"] == workload:
                return row["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------ CUDA arm
def main():
    args = parse_args()
    # stdout carries exactly ONE line, the JSON: native libraries write there too (NCCL prints its version at communicator
    # creation whatever NCCL_DEBUG says on some boxes), so file descriptor 1 is pointed at stderr for the whole run and
    # the JSON line goes to a private duplicate of the original stdout (emit()).
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = make_workload(args.workload, B=args.batch, H=args.horizon, distinct_lengthscales=args.distinct_lengthscales)
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    from rl_gp_mpc import GpMpcController, _cabi
    from rl_gp_mpc.parallel import allgather_costs, shard_bounds
    from rl_gp_mpc.config_classes.actions_config import ActionsConfig
    from rl_gp_mpc.config_classes.controller_config import ControllerConfig
    from rl_gp_mpc.config_classes.model_config import ModelConfig
    from rl_gp_mpc.config_classes.observation_config import ObservationConfig
    from rl_gp_mpc.config_classes.reward_config import RewardConfig
    from rl_gp_mpc.config_classes.total_config import Config

    E, Na, H, B, N, D = cfg["E"], cfg["Na"], cfg["H"], cfg["B"], cfg["N"], cfg["D"]
    # ---- shard the candidate batch (contiguous slices; the training block is replicated, every rank factorises)
    per, lo, hi = shard_bounds(B, world, rank)
    Bl = hi - lo
    r = cfg["reward"]
    config = Config(
        observation_config=ObservationConfig(obs_var_norm=[cfg["obs_var"]] * E),
        reward_config=RewardConfig(target_state_norm=list(r["target_state"]), weight_state=list(r["weight_state"]),
                                   weight_state_terminal=list(r["weight_state_terminal"]),
                                   target_action_norm=list(r["target_action"]), weight_action=list(r["weight_action"]),
                                   exploration_factor=r["exploration_factor"]),
        actions_config=ActionsConfig(), controller_config=ControllerConfig(len_horizon=H),
        model_config=ModelConfig(gp_init={"noise_covar.noise": list(cfg["noise"]),
                                          "base_kernel.lengthscale": [list(v) for v in cfg["lengthscale"]],
                                          "outputscale": list(cfg["outputscale"])},
                                 min_std_noise=1e-4, max_std_noise=1.0, min_outputscale=1e-6, max_outputscale=10.0,
                                 min_lengthscale=1e-3, max_lengthscale=1e3))
    ctrl = GpMpcController(-np.ones(E), np.ones(E), -np.ones(Na), np.ones(Na), config, device=dev)
    tm = ctrl.transition_model
    t0 = time.perf_counter()
    tm.prepare_inference(torch.as_tensor(cfg["x"]), torch.as_tensor(cfg["y"]))
    torch.cuda.synchronize()
    prepare_ms_first = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    tm.prepare_inference(torch.as_tensor(cfg["x"]), torch.as_tensor(cfg["y"]))
    torch.cuda.synchronize()
    prepare_ms = (time.perf_counter() - t0) * 1e3
    # one-point append (gpmpc_append): factorise N-1 points, time adding the N-th, then refactorise for the run
    tm.prepare_inference(torch.as_tensor(cfg["x"][:-1]), torch.as_tensor(cfg["y"][:-1]))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tm.prepare_inference(torch.as_tensor(cfg["x"]), torch.as_tensor(cfg["y"]))
    torch.cuda.synchronize()
    append_ms = (time.perf_counter() - t0) * 1e3 if tm.last_prepare_mode == "append" else None
    tm.incremental_updates = False
    tm.prepare_inference(torch.as_tensor(cfg["x"]), torch.as_tensor(cfg["y"]))
    tm.incremental_updates = True
    eng = tm.engine
    eng.enable_timing(True)
    obs_mu, obs_var = torch.as_tensor(cfg["mu0"]), torch.as_tensor(cfg["Sigma0"])
    actions_host = torch.as_tensor(cfg["actions"][lo:hi].reshape(Bl, H * Na)).pin_memory()
    actions_dev = actions_host.to(dev)
    mu_dev, var_dev = obs_mu.to(dev), obs_var.to(dev)
    flush = torch.empty(32 * 1024 * 1024, dtype=torch.float64, device=dev)   # 256 MB > 126 MB L2
    costs_all = torch.empty(per * world, dtype=torch.float64, device=dev)
    out = {"cost": costs_all[rank * per: rank * per + Bl]}                   # all-gather in place (no copy)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(need_grad):
        ctrl._bind_cost()
        eng.rollout(actions_dev, mu_dev, var_dev, H, need_grad=need_grad, need_traj=False, out=out)
        if dist is not None:
            allgather_costs(dist, costs_all, per, rank)

    def step_e2e():
        a = actions_host.to(dev, non_blocking=True)
        costs, grads = ctrl.compute_mean_lcb_trajectory_batch(a, obs_mu, obs_var, need_grad=True)
        if dist is not None:
            costs_all[rank * per: rank * per + Bl].copy_(costs)
            allgather_costs(dist, costs_all, per, rank)
        return costs.cpu(), grads.cpu()

    def timed(fn, steps, kernel_times=None):
        """K steps between barriers + syncs; device time by CUDA events on the launching stream; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
            if kernel_times is not None:
                torch.cuda.current_stream().synchronize()
                kernel_times.append((eng.last_rollout_ms(), eng.last_backward_ms()))
            flush.zero_()   # 256 MB write: evicts L2 between timed iterations (0.04 ms, inside the timed region)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    launches0 = None
    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        step_device(True)
        flush.zero_()
    step_device(False)
    step_e2e()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # ---- timed: fwd+grad on device-resident inputs (the `value`)
    ktimes = []
    launches0 = eng.launch_count()
    flush.zero_()
    ms_total = timed(lambda: step_device(True), args.steps, ktimes)
    launches = eng.launch_count() - launches0
    flush.zero_()
    ms_fwd = timed(lambda: step_device(False), args.steps)
    flush.zero_()
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if sampler else None
    preds = B * H
    value = preds * args.steps / (ms_total * 1e-3)
    if rank == 0:
        fwd_ms = float(np.mean([k[0] for k in ktimes]))
        bwd_ms = float(np.mean([k[1] for k in ktimes]))
        hbm_peak, peak_src = 6650.0, "fallback"
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm_peak, peak_src = float(mp["hbm_gbs"]), "measured"
        except Exception:
            pass
        b_alg = algorithmic_bytes_per_prediction(E, D, N, 8)
        f_alg = algorithmic_flops_per_prediction(E, D, N)
        preds_rank0 = Bl * H
        achieved = preds_rank0 * b_alg / (fwd_ms * 1e-3) / 1e9
        fp64_peak = _cabi.measure_fp64_peak(local_rank)
        uniform = eng.uses_uniform_path()
        kname = ("gpmpc::uniform_fwd_kernel<%d> (one exp per (i,j) for all output pairs)" % E) if uniform else \
            ("gpmpc::rollout_kernel<%d,true> (per-pair sweep + forward-mode Jacobian records)" % E)
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": ncu_traffic("uniform_fwd" if uniform else "rollout_kernel", "%s B=%d H=%d" % (cfg["name"], Bl, H)),
                    "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)",
                    "kernel": "%s, %.2f ms/launch for %d predictions" % (kname, fwd_ms, preds_rank0),
                    "reverse_sweep_kernel": ("gpmpc::uniform_bwd_kernel<%d> (adjoint-weighted upper-triangle N^2 sweep), "
                                             "%.2f ms/launch" % (E, bwd_ms)) if uniform else
                    ("gpmpc::backward_kernel<%d> (small-matrix algebra on records), %.2f ms/launch" % (E, bwd_ms)),
                    "algorithmic_bytes_per_prediction": b_alg,
                    "note": "algorithmic bytes (SURVEY 8(d): 8*(E N^2 + E N + N D) per prediction) are served from L2/L1/"
                            "shared memory -- the training block is shared by all candidates (`traffic` = DRAM bytes per launch "
                            "from ncu, mostly the per-step records and outputs) -- so frac>1 is expected; the binding roofline is float64 FMA "
                            "throughput (fp64 block)",
                    "fp64": fp64_report(uniform, E, N, preds_rank0, fwd_ms, bwd_ms, f_alg, fp64_peak)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(cfg, args),
                "forward_only": {"value": preds * args.steps / (ms_fwd * 1e-3), "unit": UNIT,
                                 "ms_per_step": ms_fwd / args.steps},
                "kernel_path": "uniform (all GPs share their hyper-parameters)" if uniform else "general (per-pair)",
                "kernel_ms": {"rollout_fwd": fwd_ms, "reverse_sweep": bwd_ms},
                "prepare_ms": {"first_call": prepare_ms_first, "steady": prepare_ms, "append_one_point": append_ms,
                               "what": "Gram + Cholesky + iK + beta for %d GPs, N=%d (once per control step); "
                                       "append_one_point = gpmpc_append through prepare_inference when the memory grew by one" % (E, N)},
                "e2e": {"value": preds * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": int(actions_host.numel() * 8 * world),
                        "d2h_bytes_per_step": int((Bl + Bl * H * Na) * 8 * world),
                        "api": "GpMpcController.compute_mean_lcb_trajectory_batch (pinned host actions in, costs+grads out)"},
                "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks}
        if not args.no_cpu_baseline:
            rate, dt = cpu_port_rate(cfg, min(16, B), H)      # ~10-15 s of host work at the headline shape
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "%d candidates x %d horizon steps, objective+autograd gradient, %.1f s "
                                              "(oracle/gpmpc_oracle.py, float64)" % (min(16, B), H, dt)}
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
