#!/usr/bin/env python
"""bench.py -- throughput of the Rabacus hot path on B200(s).

    python bench.py --gpus N --steps K --warmup W            # CUDA arm
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm

Metric (BASELINE.json): SphereBgnd cell*ray*frequency evaluations per second.
Workload (BASELINE.json configs[3]): SphereBgnd, HM12 z=3 background, 4096
shells x 512 frequencies x 256 polar rays, fixed T (rabacus_b200.configs.
c4_sphere_bgnd).  A "step" is one Jacobi sweep of the solver over the whole
sphere: pack shells -> ray-segment columns -> fused tau/exp/quadrature ->
per-cell equilibrium solve -> convergence test.

* ``value``: K sweeps with the problem resident in HBM, timed with CUDA events
  on the stream the kernels run on; for N > 1 the cells are dealt cyclically over
  the ranks and every sweep ends with one NCCL all-gather of the updated state
  (strong scaling: total work fixed; time = max over ranks).
* ``e2e``: the same metric through the reference-facing C-ABI entry point
  ``rb_sphere_bgnd_solve`` with HOST buffers: upload, optically-thin start,
  sweeps to convergence (tol 1e-4), download; evaluations of the whole solve /
  wall time of the call.
* ``roofline``: the dominant kernel against the FP64 pipe, peak measured in this
  run by a DFMA micro-benchmark (MEASURED_PEAKS.json has no FP64 number).
* ``cpu_baseline``: the C restatement of the reference (oracle/, "port": the
  Fortran cannot be built here) on this box's host cores, one sweep of the same
  workload.  ``--impl reference`` times that arm alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SphereBgnd cell*ray*freq evals/s"
UNIT = "evals/s"

# algorithmic FP64-pipe operations (DESIGN.md "Roofline"):
#   frequency kernel, per (cell, ray, nu): 3 (tau = sum_X tau_X ra_X) + 8 (exp: scale,
#   round, reduce, 4 polynomial, table merge) + 1 (A(nu) += w_ray a); the six rate
#   accumulations are per (cell, nu), 1/Nmu of that, and are not counted
#   column kernel, per shell edge of a ray pair: 1 (E^2 - p^2) + 5 (sqrt: MUFU seed + one
#   third-order correction) + 3 (column FMAs, summed by parts: P += s_k (c_k - c_{k-1}))
OPS_PER_EVAL = 12
OPS_PER_SEGMENT = 9


def workload(args):
    from rabacus_b200 import configs
    return configs.c4_sphere_bgnd(args.Nl, args.Nnu, args.Nmu, find_Teq=bool(args.find_Teq))


def workload_name(args):
    return ("SphereBgnd HM12 z=3, %d shells x %d freqs x %d rays, %s, rec_meth=fixed"
            % (args.Nl, args.Nnu, args.Nmu, "find_Teq" if args.find_Teq else "fixed T=1e4K"))


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (NVML in
    process, every ~2 ms; nvidia-smi subprocess as a fallback)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
            ("sw_power_cap", 0x4))

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.stop_flag = index, False
        self.sm, self.max_mhz, self.reasons, self.n = [], None, set(), 0
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        p = self.nvml
        self.sm.append(int(p.nvmlDeviceGetClockInfo(self.h, p.NVML_CLOCK_SM)))
        try:
            r = p.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = p.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in self.BITS:
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True,
                             timeout=5).stdout.strip()
        c = [v.strip() for v in out.split(",")]
        if c and c[0].isdigit():
            self.sm.append(int(c[0]))
            self.max_mhz = int(c[1]) if c[1].isdigit() else self.max_mhz
            for i, (name, _) in enumerate(self.BITS):
                if len(c) > 2 + i and c[2 + i].lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                    time.sleep(0.002)
                else:
                    self._sample_smi()
                    time.sleep(0.05)
                self.n += 1
            except Exception:
                time.sleep(0.05)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def segments_per_sweep(cfg, Nmu):
    """Exact count of ray segments whose chord length the column kernel evaluates
    (one sqrt each): sum over cells and |mu| pairs of (m+1)."""
    import numpy as np
    from rabacus_b200 import rabacus_fc
    E = cfg["Edges"][::-1].copy()  # decreasing
    r = 0.5 * (E[:-1] + E[1:])
    mu, _ = rabacus_fc.p_quadrature_rule(Nmu)
    total = 0
    asc = E[::-1]
    for ip in range((Nmu + 1) // 2):
        p = r * np.sqrt(1.0 - mu[ip] * mu[ip])
        # m = (first i>=1 with E_i <= p) - 1  == (number of edges > p) - 1
        n_gt = E.size - np.searchsorted(asc, p, side="right")
        total += int(np.sum(n_gt))
    return total


def cpu_sweep_seconds(cfg, args):
    """One Jacobi sweep of the workload on the host cores with the oracle (the C
    restatement of the reference, OpenMP over shells like sphere_bgnd.f90:753-913)."""
    from oracle import oracle
    oracle.build()
    t0 = time.perf_counter()
    oracle.sphere_bgnd_solve(cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["Nmu"],
                             cfg["E_eV"], cfg["shape"], i_find_Teq=1 if cfg["find_Teq"] else 0,
                             z=cfg["z"], tol=1e-4, i_thin=1)
    t_thin = time.perf_counter() - t0
    t0 = time.perf_counter()
    res = oracle.sphere_bgnd_solve(cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["Nmu"],
                                   cfg["E_eV"], cfg["shape"],
                                   i_find_Teq=1 if cfg["find_Teq"] else 0, z=cfg["z"], tol=1e-4,
                                   order=oracle.ORDER_JACOBI, max_sweeps=1)
    t = time.perf_counter() - t0 - t_thin
    return max(t, 1e-9), oracle.num_threads(), res


def run_reference(args, rank, world, out):
    """CPU arm: the reference's algorithm on the host cores (oracle port)."""
    if rank != 0:
        return
    cfg = workload(args)
    evals = args.Nl * args.Nmu * args.Nnu
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_sweep_seconds(cfg, args)
    ts = []
    cores = 1
    for _ in range(args.steps):
        t, cores, _ = cpu_sweep_seconds(cfg, args)
        ts.append(t)
    total = sum(ts)
    value = evals * len(ts) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "step": "one Jacobi sweep (thin start excluded)",
                   "sweep_order": "jacobi"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d full sweep(s) of the workload, OMP_NUM_THREADS=%d, C "
                                   "restatement of the Fortran with the tangent-shell search "
                                   "hoisted and one exp per evaluation (same arithmetic)"
                                   % (len(ts), cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=out, flush=True)


def _claim_stdout():
    """stdout must carry exactly ONE JSON line: everything else (NCCL's version
    banner, library chatter) is sent to stderr; returns a handle on the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--Nl", type=int, default=4096)
    ap.add_argument("--Nnu", type=int, default=512)
    ap.add_argument("--Nmu", type=int, default=256)
    ap.add_argument("--find_Teq", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how the updated fractions reach the other ranks each sweep")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    out = _claim_stdout()

    if args.impl == "reference":
        run_reference(args, rank, world, out)
        return

    import numpy as np
    import torch
    from rabacus_b200 import _lib, configs, rabacus_fc
    from rabacus_b200.plan import Plan

    if _lib.device_count() < 1:
        raise SystemExit("bench.py --impl cuda needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W = max(3, args.warmup)
    K = max(1, args.steps)

    cfg = workload(args)
    evals_sweep = args.Nl * args.Nmu * args.Nnu
    plan = Plan(configs.GEOM_IDS["sphere_bgnd"], 1, args.Nl, args.Nnu, args.Nmu,
                find_Teq=bool(args.find_Teq), fixed_fcA=1.0, tol=1e-4, device=local_rank,
                max_sweeps=-(10 ** 9))  # fixed-work mode: keep sweeping past convergence
    plan.set_problem(0, cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"],
                     cfg["shape"], cfg["z"], 0.0)
    plan.upload()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    flush_bytes = 256 << 20
    launches0 = 0

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if world == 1 or args.exchange == "peer":
        # N > 1: the fractions travel as remote stores from the cell-solver kernel into
        # every rank's staging buffer + a flag barrier in the finish kernel (no NCCL on
        # the per-sweep path); the loop below is the same C call on every rank.
        if world > 1:
            plan.connect_peers(rank, world)
        plan.thin()
        plan.bench_steps(W, flush_bytes)
        info0 = plan.last_info()
        launches0 = info0["n_launches"]
        barrier()
        if sampler:
            sampler.start()
        ms = plan.bench_steps(K, flush_bytes)
        barrier()
        info = plan.last_info()
    else:
        send, recv = plan.shard_buffers(world)
        ext = torch.cuda.ExternalStream(plan.stream_ptr)
        flush = torch.empty(flush_bytes, dtype=torch.uint8, device="cuda")
        with torch.cuda.stream(ext):
            plan.thin()
            for _ in range(W):
                plan.sharded_sweep(rank, world, send, recv)
            info0 = plan.last_info()
            launches0 = info0["n_launches"]
            barrier()
            if sampler:
                sampler.start()
            ms = 0.0
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            for s in range(K):
                flush.fill_(s & 0xff)
                e0.record(ext)
                plan.sharded_sweep(rank, world, send, recv)
                e1.record(ext)
                e1.synchronize()
                ms += e0.elapsed_time(e1)
            barrier()
        info = plan.last_info()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = info["n_launches"] - launches0
    value = evals_sweep * K / (ms * 1e-3)

    # stage times of the K timed sweeps (warm-up stage times subtracted)
    plan_all = info

    # ---------------------------------------------------------------- e2e
    e2e = None
    solve_ms = None
    iters = None
    if not args.no_e2e:
        reps = max(1, min(K, 3))
        if world == 1:
            def one_call():
                T = cfg["T"].copy()
                rabacus_fc.sphere_bgnd.sphere_bgnd_solve(
                    cfg["Edges"].copy(), cfg["nH"].copy(), cfg["nHe"].copy(), T, args.Nmu,
                    cfg["E_eV"], cfg["shape"], 1, 1.0, 1, 1, args.find_Teq, 0, 1.0, 1.0, 1.0,
                    cfg["z"], 0.0, 1e-4, args.Nl, args.Nnu)
                return dict(rabacus_fc.last_info)
        else:
            # the plan (buffers + peer mappings) is set up once, like the plan cache of
            # the one-shot entry point; every call stages, uploads, solves, downloads
            p2 = Plan(configs.GEOM_IDS["sphere_bgnd"], 1, args.Nl, args.Nnu, args.Nmu,
                      find_Teq=bool(args.find_Teq), fixed_fcA=1.0, tol=1e-4, device=local_rank)
            if args.exchange == "peer":
                p2.connect_peers(rank, world)

            def one_call():
                p2.set_problem(0, cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"],
                               cfg["shape"], cfg["z"], 0.0)
                p2.upload()
                n = p2.solve_cell_sharded(rank, world)
                p2.get_problem(0)
                inf = p2.last_info()
                inf["iters"] = n
                return inf
        one_call()  # warm (module load, allocations)
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            inf = one_call()
        barrier()
        dt = (time.perf_counter() - t0) / reps
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        iters = int(inf["iters"])
        solve_ms = dt * 1e3
        h2d = 8 * (args.Nl + 1 + 3 * args.Nl + args.Nnu * 9 + 6 + 2)
        d2h = 8 * 12 * args.Nl
        e2e = {"value": evals_sweep * iters / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "step": "one rb_sphere_bgnd_solve call (upload, thin "
               "start, %d sweeps to tol=1e-4, download)" % iters, "solve_wall_ms": solve_ms,
               "sweeps": iters}

    # the workload's second flavour ("both fixed T and find_Teq"): wall time of
    # one converged equilibrium-temperature solve through the same entry point
    solve_teq = None
    if e2e is not None and world == 1 and not args.find_Teq:
        def teq_call():
            T = cfg["T"].copy()
            rabacus_fc.sphere_bgnd.sphere_bgnd_solve(
                cfg["Edges"].copy(), cfg["nH"].copy(), cfg["nHe"].copy(), T, args.Nmu,
                cfg["E_eV"], cfg["shape"], 1, 1.0, 1, 1, 1, 0, 1.0, 1.0, 1.0,
                cfg["z"], 0.0, 1e-4, args.Nl, args.Nnu)
            return dict(rabacus_fc.last_info)
        teq_call()
        t0 = time.perf_counter()
        inf = teq_call()
        solve_teq = {"sweeps_to_converge": int(inf["iters"]),
                     "wall_ms": (time.perf_counter() - t0) * 1e3}

    if sampler:  # sampled over the kernel-timed region and the e2e solves
        sampler.stop_flag = True
        sampler.join(timeout=2)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- roofline
    import ctypes as C
    peak = {}
    for kind, name in ((0, "dfma"), (1, "exp"), (2, "sqrt"), (3, "pow")):
        v = C.c_double(0.0)
        _lib.check(_lib.lib().rb_microbench(kind, local_rank, C.byref(v)))
        peak[name] = v.value
    n_seg = segments_per_sweep(cfg, args.Nmu)
    # stage times of the K timed sweeps only (warm-up sweeps subtracted)
    ms_cols = (plan_all["ms_columns"] - info0["ms_columns"]) / K
    ms_nu = (plan_all["ms_nu"] - info0["ms_nu"]) / K
    ms_cell = (plan_all["ms_cell"] - info0["ms_cell"]) / K
    if world > 1:
        n_seg_local, ev_local = n_seg / world, evals_sweep / world
    else:
        n_seg_local, ev_local = n_seg, evals_sweep
    peak_tflops = 2.0 * peak["dfma"] / 1e12
    kern = {
        "k_nu_rates": {"ms": ms_nu, "alg_ops": ev_local * OPS_PER_EVAL},
        "k_sphere_columns": {"ms": ms_cols, "alg_ops": n_seg_local * OPS_PER_SEGMENT},
    }
    for k in kern.values():
        k["tflops"] = 2.0 * k["alg_ops"] / (k["ms"] * 1e-3) / 1e12 if k["ms"] > 0 else 0.0
        k["frac"] = k["tflops"] / peak_tflops if peak_tflops > 0 else 0.0
    dom = max(kern, key=lambda n: kern[n]["ms"])
    traffic = None  # DRAM bytes per launch of that kernel from the committed ncu capture
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f)["bytes_per_launch"].get(dom) if world == 1 else None
    except Exception:
        traffic = None
    roofline = {
        "bound": "fp64", "kernel": dom, "achieved": kern[dom]["tflops"], "peak": peak_tflops,
        "unit": "TFLOP/s", "frac": kern[dom]["frac"], "traffic": traffic,
        "traffic_source": "profiles/ncu_traffic.json (ncu --set full, dram__bytes_read.sum + "
                          "dram__bytes_write.sum per launch; the 33.5 MB ray buffer stays in L2)",
        "peak_source": "measured in this run: DFMA micro-benchmark (rb_microbench), "
                       "MEASURED_PEAKS.json has no FP64 figure",
        "ms_per_launch": kern[dom]["ms"],
        "kernels": {n: {"ms": round(v["ms"], 4), "tflops": round(v["tflops"], 3),
                        "frac": round(v["frac"], 4)} for n, v in kern.items()},
        "ms_cell_solve": round(ms_cell, 4),
        "ms_finish": round((plan_all["ms_finish"] - info0["ms_finish"]) / K, 4),
        "step_frac_of_fp64_peak": (2.0 * (ev_local * OPS_PER_EVAL + n_seg_local * OPS_PER_SEGMENT)
                                   / (ms / K * 1e-3) / 1e12) / peak_tflops,
        "microbench_ops_per_s": peak,
        "segments_per_sweep": n_seg,
    }

    # ---------------------------------------------------------------- CPU baseline
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            t, cores, _ = cpu_sweep_seconds(cfg, args)
            cpu = {"value": evals_sweep / t, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "1 full Jacobi sweep of the workload (%.1f s), C restatement of the "
                             "Fortran (oracle/), OpenMP over shells, tangent-shell search hoisted"
                             % t}
        except Exception as e:  # the CPU leg must never break the GPU line
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                   "sample": "failed: %r" % (e,)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args),
                   "step": "one Jacobi sweep over all cells incl. convergence test",
                   "parallelism": "cells dealt cyclically over %d GPU(s)%s"
                                  % (world, "" if world == 1 else
                                     (", fractions exchanged by remote stores over NVLink peer "
                                      "memory + flag barrier (no collective per sweep)"
                                      if args.exchange == "peer" else
                                      ", NCCL all-gather of 18 state fields per sweep")),
                   "l2": "256 MiB scratch overwritten between steps (not timed); working set "
                         "per sweep ~25 MB of column data, compute bound"},
        "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        "clocks": sampler.summary() if sampler else None,
        "solve": {"sweeps_to_converge": iters, "wall_ms": solve_ms},
        "solve_find_Teq": solve_teq,
    }
    print(json.dumps(line), file=out, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
