#!/usr/bin/env python
"""bench.py -- collocation-points/sec of one loss + gradient evaluation (BASELINE.json metric).

A "step" is one pass of the hot path (pinn_loss_grad: fused forward-taps / residual /
reverse sweep + gradient reduction [+ NCCL allreduce at N>1]) over the workload's point
sets.  Workload at N=1: BASELINE.json configs[1] -- 2-D Poisson on [0,1]^2, 4x64 tanh MLP,
GridTraining with 128^2 collocation points, fp32.  At N>1 every rank holds a 128^2 shard of
a 128 x (128 N) grid (weak scaling) and the gradient is all-reduced once per step.

  python bench.py --gpus N --steps K --warmup W            # our engine
  python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference

Prints ONE JSON line (see the contract in the task statement).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "collocation-points/sec (loss+grad)"
UNIT = "points/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """Samples SM clock, power and throttle reasons during the timed region (NVML, ~2 ms period;
    falls back to nvidia-smi)."""

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
        while not self._stop.is_set():
            try:
                clk = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((float(clk), float(mx), {k for k, b in bits.items() if rs & b}))
            except Exception:
                pass
            self._stop.wait(0.002)

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                p = [s.strip() for s in out.strip().split(",")]
                if len(p) >= 6:
                    self.rows.append((float(p[0]), float(p[1]), {n for n, v in zip(names, p[2:6]) if v.lower().startswith("active")}))
            except Exception:
                pass
            self._stop.wait(0.05)

    def _run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = [r[0] for r in self.rows]
        reasons = set()
        for r in self.rows:
            reasons |= r[2]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(r[1] for r in self.rows) if self.rows else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def build_workload(world: int, n: int = 128):
    from neuralpde_jl_b200 import configs
    from neuralpde_jl_b200.strategies import GridTraining
    cfg = configs.config2(n=n)
    if world > 1:
        cfg.strategy = GridTraining([1.0 / (n - 1), 1.0 / (n * world - 1)])
    return cfg


def cpu_reference_eval(cfg, theta64, sets, threads: int, reps: int):
    """The reference algorithm on the host: finite-difference stencils (K forward passes per
    PDE term), mean(abs2), reverse-mode gradient, float64, all cores (oracle/reference.py)."""
    import torch
    from oracle import reference as R
    torch.set_num_threads(threads)
    prob = R.Problem(cfg.pde_system, cfg.chain_specs(), param_estim=cfg.param_estim, derivative="fd")
    n_pde = len(cfg.pde_system.eqs)
    ps, bs = sets[:n_pde], sets[n_pde:]
    times, L = [], None
    for _ in range(reps):
        t0 = time.perf_counter()
        L, _, _ = prob.loss_and_grad(theta64, ps, bs)
        times.append(time.perf_counter() - t0)
    return L, times


def best_thread_count(cfg, theta64, sets, cores: int) -> int:
    """The reference side gets the thread count that serves it best (oversubscribing small GEMMs on a
    many-core host is slower than using fewer threads)."""
    cands = sorted({c for c in (8, 16, 32, 64, cores) if c <= cores})
    best, best_t = cands[0], float("inf")
    for c in cands:
        cpu_reference_eval(cfg, theta64, sets, c, 1)
        _, t = cpu_reference_eval(cfg, theta64, sets, c, 2)
        if min(t) < best_t:
            best, best_t = c, min(t)
    return best


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    import torch  # noqa: F401
    from oracle import reference as R
    cfg = build_workload(world, args.n)
    sys_ = cfg.pde_system
    ps, bs = R.generate_training_sets(sys_.domain, cfg.strategy.dx, sys_.eqs, sys_.bcs, sys_.ivs, sys_.dvs)
    n_pts = sum(p.shape[1] for p in ps)
    theta = cfg.init_params(np.float64)
    cores = best_thread_count(cfg, theta, ps + bs, os.cpu_count() or 1)
    _, tw = cpu_reference_eval(cfg, theta, ps + bs, cores, 1)          # warm-up, also sizes the sample
    budget = 150.0                                                     # seconds for the K timed steps
    frac = min(1.0, budget / max(args.steps * tw[0], 1e-9))
    if frac < 1.0:                                                     # bounded sample: leading fraction of every set
        ps = [p[:, :max(1, int(np.ceil(frac * p.shape[1])))] for p in ps]
        bs = [b[:, :max(1, int(np.ceil(frac * b.shape[1])))] for b in bs]
    n_pts_full = n_pts
    n_pts = sum(p.shape[1] for p in ps)
    for _ in range(max(0, min(args.warmup, 3) - 2)):
        cpu_reference_eval(cfg, theta, ps + bs, cores, 1)
    L, times = cpu_reference_eval(cfg, theta, ps + bs, cores, args.steps)
    total = float(np.sum(times))
    val = n_pts * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "2-D Poisson, 4x64 tanh MLP, GridTraining %dx%d points (BASELINE configs[1])"
                               % (args.n, args.n * world), "n_pde_points": n_pts, "loss": L},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d loss+grad evaluations over %d of the workload's %d PDE points (+ the same "
                                   "fraction of every bc set); CPU restatement of the reference algorithm (FD "
                                   "stencils, PyTorch-CPU float64, all cores), not Julia"
                                   % (args.steps, n_pts, n_pts_full)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist
    import neuralpde_jl_b200 as npde

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = build_workload(world, args.n)
    dtype = np.float32
    disc = cfg.discretization(dtype=dtype, mode=args.mode, device=local_rank)
    rep = npde.symbolic_discretize(cfg.pde_system, disc, rank=rank, world=world)
    eng = rep.engine
    n_pde = len(cfg.pde_system.eqs)
    n_pts_global = sum(rep.point_sets[i].shape[1] for i in range(n_pde))
    if world > 1:
        uid = [npde.Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(uid[0], rank, world)

    theta_h = rep.flat_init_params
    n_theta, n_terms = eng.n_theta, eng.n_terms
    theta_d = torch.from_numpy(theta_h).to(dev)
    grad_d = torch.empty(n_theta, dtype=torch.float32, device=dev)
    terms_d = torch.empty(n_terms, dtype=torch.float32, device=dev)
    total_d = torch.empty(1, dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 256 MiB > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        eng.loss_grad_device(theta_d, grad_d, terms_d, total_d, None, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        step()
    barrier()

    # ---- device-resident timing: K steps, each bracketed by CUDA events, L2 flushed between steps ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in ev:
        flush.zero_()
        a.record()
        step()
        b.record()
    barrier()
    launches = eng.launch_count() - launches0
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    t_total = float(ms.sum()) * 1e-3
    loss_val = float(total_d.item())

    # ---- main-kernel duration (for the roofline), events inside the library around the fused kernel ----
    eng.set_timing(True)
    kms = []
    for _ in range(min(args.steps, 20)):
        flush.zero_()
        step()
        kms.append(eng.last_kernel_ms())
    eng.set_timing(False)
    kernel_ms = float(np.mean(kms))

    # ---- end to end through host buffers (H2D theta, D2H grad + losses inside the timed region) ----
    for _ in range(3):
        eng.loss_grad_host(theta_h, None, True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tot_h, _, _ = eng.loss_grad_host(theta_h, None, True)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        tt = torch.tensor([t_total, t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_total, t_e2e = float(tt[0]), float(tt[1])

    if rank == 0:
        pk, pk_kind = peaks()
        flops = eng.flops_per_eval() * (1 if world == 1 else 1)      # per-rank launch
        achieved = flops / (kernel_ms * 1e-3) / 1e12
        value = n_pts_global * args.steps / t_total
        # DRAM traffic of the dominant kernel: measured once per round with `ncu --set full` (profiles/), not re-measured here
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_ncu_traffic.json")) as fh:
                ent = json.load(fh).get("%s_n%d" % (args.mode, args.n))
            if ent and world == 1:
                traffic, traffic_src = int(ent["dram_bytes_read"]) + int(ent["dram_bytes_write"]), ent["source"]
        except (OSError, ValueError, KeyError):
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "2-D Poisson on [0,1]^2, 4x64 tanh MLP, GridTraining %dx%d points, fp32 "
                                   "(BASELINE configs[1]%s)" % (args.n, args.n * world,
                                                                 "" if world == 1 else "; %d-point shard per rank" % (args.n * args.n)),
                       "n_pde_points": n_pts_global, "n_bc_points": sum(p.shape[1] for p in rep.point_sets[n_pde:]),
                       "mode": args.mode, "l2": "flushed between timed steps (256 MiB memset)",
                       "parallelism": "dp%d" % world, "loss": loss_val, "n_theta": n_theta},
            "clocks": clocks,
            "e2e": {"value": n_pts_global * args.steps / t_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": int(n_theta * 4), "d2h_bytes_per_step": int((n_theta + n_terms + 1) * 4),
                    "ms_per_step": 1e3 * t_e2e / args.steps, "api": "pinn_loss_grad_host"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": achieved / pk["bf16_tflops"], "traffic": traffic, "traffic_unit": "bytes (DRAM read + write per launch, ncu)",
                         "traffic_source": traffic_src, "peak_source": pk_kind,
                         "kernel": "ffma_loss_grad_kernel" if args.mode == "ffma" else "tc_loss_grad_kernel",
                         "arithmetic": {"ffma": "fp32 FMA on CUDA cores", "tc_bf16": "tcgen05 bf16 x bf16 -> fp32",
                                        "tc_split": "tcgen05 split-bf16 (3 MMAs per product in the forward sweep)"}[args.mode],
                         "kernel_ms": kernel_ms, "flops_per_launch": flops,
                         "note": "algorithmic FLOPs 6*C*S per point (SURVEY 8(d)) / fused-kernel duration"},
        }
        if world == 1 and not args.no_alt_modes:
            # the other arithmetic modes on the same workload (device-resident, kernel + reduction), for context
            alt = {}
            for m in ("ffma", "tc_bf16", "tc_split"):
                if m == args.mode:
                    continue
                d2 = cfg.discretization(dtype=dtype, mode=m, device=local_rank)
                r2 = npde.symbolic_discretize(cfg.pde_system, d2)
                for _ in range(3):
                    r2.engine.loss_grad_device(theta_d, grad_d, terms_d, total_d, None, stream)
                e2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
                for a, b in e2:
                    flush.zero_()
                    a.record()
                    r2.engine.loss_grad_device(theta_d, grad_d, terms_d, total_d, None, stream)
                    b.record()
                torch.cuda.synchronize()
                ms2 = float(np.mean([a.elapsed_time(b) for a, b in e2]))
                alt[m] = {"ms_per_step": ms2, "value": n_pts_global / (ms2 * 1e-3), "loss": float(total_d.item())}
                r2.engine.close()
            line["config"]["other_modes"] = alt
            # BASELINE configs[2] (Burgers, 5x128 MLP, 65 536 stochastic points) on the 128-wide tcgen05 path, for context:
            # a parity-test configuration, not the headline workload
            try:
                from neuralpde_jl_b200 import configs as _cfgs
                c3 = _cfgs.config3()
                r3 = npde.symbolic_discretize(c3.pde_system, c3.discretization(dtype=dtype, mode="tc_bf16", device=local_rank))
                th3 = torch.from_numpy(r3.flat_init_params).to(theta_d.device)
                g3 = torch.empty_like(th3)
                t3 = torch.empty(r3.engine.n_terms, dtype=th3.dtype, device=th3.device)
                r3.loss_functions.full_loss_function(r3.flat_init_params)       # draws + uploads the first sample
                for _ in range(3):
                    r3.engine.loss_grad_device(th3, g3, t3, total_d, None, stream)
                e3 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
                for a, b in e3:
                    flush.zero_()
                    a.record()
                    r3.engine.loss_grad_device(th3, g3, t3, total_d, None, stream)
                    b.record()
                torch.cuda.synchronize()
                ms3 = float(np.mean([a.elapsed_time(b) for a, b in e3]))
                fl3 = float(r3.engine.flops_per_eval())
                line["config"]["other_configs"] = {"cfg3_burgers_5x128_tc_bf16": {
                    "ms_per_step": ms3, "value": c3.n_pde_points / (ms3 * 1e-3), "unit": UNIT,
                    "tflops_algorithmic": fl3 / (ms3 * 1e-3) / 1e12, "frac_of_bf16_peak": fl3 / (ms3 * 1e-3) / 1e12 / pk["bf16_tflops"],
                    "kernel": "tw_loss_grad_kernel", "loss": float(total_d.item())}}
                r3.engine.close()
            except Exception as ex:      # context only: never fail the headline line
                line["config"]["other_configs"] = {"cfg3_burgers_5x128_tc_bf16": {"error": str(ex)[:200]}}
        if world == 1 and not args.no_cpu_baseline:
            sets = rep.point_sets[:n_pde + len(cfg.pde_system.bcs)]
            cores = best_thread_count(cfg, theta_h.astype(np.float64), sets, os.cpu_count() or 1)
            reps = 5
            cpu_reference_eval(cfg, theta_h.astype(np.float64), sets, cores, 1)
            Lc, times = cpu_reference_eval(cfg, theta_h.astype(np.float64), sets, cores, reps)
            line["cpu_baseline"] = {
                "value": n_pts_global / float(np.median(times)), "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "%d full loss+grad evaluations of the same workload (median); CPU restatement of the "
                          "reference algorithm (FD stencils, PyTorch-CPU float64, best of 8/16/32/64/all threads), not "
                          "Julia" % reps,
                "host_cores": os.cpu_count(), "loss": Lc, "loss_rel_err_engine": abs(loss_val - Lc) / abs(Lc)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("PINN_BENCH_MODE", "tc_split"), choices=["ffma", "tc_bf16", "tc_split"])
    ap.add_argument("--n", type=int, default=128, help="grid points per axis (128 = BASELINE configs[1])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-alt-modes", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: NCCL's version banner / debug output (printed to stdout when NCCL_DEBUG is
    # set in the environment) goes to stderr instead
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"      # NCCL honours NCCL_DEBUG_FILE only above the VERSION level
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
