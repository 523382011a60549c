#!/usr/bin/env python
"""bench.py -- throughput of the Rabacus hot path on B200(s).

    python bench.py --gpus N --steps K --warmup W            # CUDA arm
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm

Metric (BASELINE.json): SphereBgnd cell*ray*frequency evaluations per second.
Workload (BASELINE.json configs[3], "C4"): SphereBgnd, HM12 z=3 background, 4096
shells x 512 frequencies x 256 polar rays, fixed T (rabacus_b200.configs.
c4_sphere_bgnd).  A "step" is one Jacobi sweep of the solver over the whole
sphere: ray-segment columns -> fused tau/exp/quadrature -> per-cell equilibrium
solve -> convergence test.

* ``value``: K sweeps with the problem resident in HBM, every sweep bracketed by CUDA
  events on the stream the kernels run on.  N > 1: the cells are dealt over the ranks in
  32-cell blocks (cyclically), the new fractions travel as remote stores over NVLink peer memory and one
  flag barrier ends the sweep (no collective; ``--exchange nccl`` selects an all-gather
  instead); strong scaling: total work fixed, time = max over ranks.  At N > 1 the timed
  sweeps are replayed as CUDA graphs, as rb_plan_solve does.
* ``e2e``: the same metric through the reference-facing C-ABI entry point
  ``rb_sphere_bgnd_solve`` with HOST buffers: upload, optically-thin start, sweeps to
  convergence (tol 1e-4), download; evaluations of the whole solve / wall time of the call.
* ``roofline``: the dominant kernel against the FP64 pipe: the implementation's FP64
  recipe against the DFMA rate measured in this run and against the architectural peak,
  the work it executes (the column kernel takes the far part of every ray from a moment
  table), SURVEY 8(d)'s library-rate model, and DRAM traffic from the committed ncu
  captures (see ``note`` / ``traffic_source`` there).
* ``cpu_baseline``: the C restatement of the reference (oracle/, "port": the Fortran
  cannot be built here), -O3 -mtune=native build, on this box's host cores.
  ``--impl reference`` times that arm alone, with all host threads.
* extra keys: ``c4_find_Teq`` (the workload's evolving-temperature flavour), ``c5`` /
  ``c5_find_Teq`` (BASELINE.json configs[4]: 8192 independent solves dealt over the
  ranks), ``zones`` (single-zone batch API at 1e6 / 1e7 zones), ``parity_vs_n1`` (N > 1:
  the sharded solve against rank 0 solving alone).
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

# algorithmic FP64-pipe operations of THIS implementation (DESIGN.md "Roofline"):
#   frequency kernel, per (cell, ray, nu): 3 (tau = sum_X tau_X ra_X) + 7 (exp: scale,
#   round, reduce, 3 polynomial, table merge) + 1 (A(nu) += w_ray a); the six rate
#   accumulations are per (cell, nu), 1/Nmu of that, and are not counted
#   column kernel, per shell edge of a ray pair: 1 (E^2 - p^2) + 5 (sqrt: MUFU seed + one
#   third-order correction) + 3 (column FMAs, summed by parts: P += s_k (c_k - c_{k-1}))
OPS_PER_EVAL = 11    # 3 (tau) + 7 (exp: 1024-entry table + cubic) + 1 (A += w a)
OPS_PER_SEGMENT = 9


def workload(args, find_Teq=None):
    from rabacus_b200 import configs
    teq = bool(args.find_Teq) if find_Teq is None else find_Teq
    return configs.c4_sphere_bgnd(args.Nl, args.Nnu, args.Nmu, find_Teq=teq)


def workload_name(args):
    return ("SphereBgnd HM12 z=3, %d shells x %d freqs x %d rays, %s, rec_meth=fixed"
            % (args.Nl, args.Nnu, args.Nmu, "find_Teq" if args.find_Teq else "fixed T=1e4K"))


def bench_config(args):
    """The SAME dict from both arms (the driver compares them)."""
    return {"workload": workload_name(args),
            "step": "one Jacobi sweep over all cells (thin start excluded)",
            "sweep_order": "jacobi",
            "l2": "GPU arm: 256 MiB scratch overwritten between steps (not timed); the sweep's "
                  "working set is ~34 MB of ray records, compute bound"}


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


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
    """Exact count of shell edges whose chord root the column kernel evaluates
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


FAR_RATIO, FAR_NT = 1.25, 25      # csrc/rb_far_coeffs.h


def far_row_slog(Nl):
    """csrc/rb_kernels.cuh far_row_slog(): every 2^slog-th far-field row is kept"""
    slog = 3
    while slog < 12 and (Nl + 1 + 8) * 32 + ((Nl >> slog) + 1) * 3 * FAR_NT * 8 > 226 * 1024:
        slog += 1
    return slog


def column_work_executed(cfg, Nmu):
    """What the far-field column kernel EXECUTES per sweep (csrc/rb_kernels.cuh
    sphere_ray_pair): shell edges walked with one sqrt each (those below the last kept
    far-field row at or above 1.25 p, plus the few edges between a kept row and the own
    shell), and far-field series evaluations (3 x 25 fused multiply-adds each)."""
    import numpy as np
    from rabacus_b200 import rabacus_fc
    E = cfg["Edges"][::-1].copy()  # decreasing, outermost first
    Nl = E.size - 1
    asc = E[::-1]
    r = 0.5 * (E[:-1] + E[1:])
    il = np.arange(Nl)
    mu, _ = rabacus_fc.p_quadrature_rule(Nmu)
    S = 1 << far_row_slog(Nl)
    walked = evals = 0
    for ip in range((Nmu + 1) // 2):
        p = r * np.sqrt(1.0 - mu[ip] * mu[ip])
        m = (E.size - np.searchsorted(asc, p, side="right")) - 1          # last edge > p
        kf = (E.size - np.searchsorted(asc, p * FAR_RATIO, side="left")) - 1   # last edge >= 1.25 p
        kf = np.minimum(kf, m)
        kc = np.where(kf >= 0, kf - kf % S, -1)
        walked += int(np.sum(m - kc))
        evals += int(np.sum(kc >= 0))
        own_far = (kc >= 0) & (il <= kc) & (il >= 1)      # own shell above the walk
        ks = (il - 1) - (il - 1) % S
        walked += int(np.sum(np.where(own_far, (il - 1) - ks + 2, 0)))   # + the two own roots
        evals += int(np.sum(own_far))
    return walked, evals


# ------------------------------------------------------------------------------ CPU legs
def cpu_sweep_seconds(cfg, sweeps=1, flavour=0):
    """`sweeps` Jacobi sweeps of cfg on the host cores with the oracle's timing build (the C
    restatement of the reference, OpenMP over shells like sphere_bgnd.f90:753-913); returns
    the wall seconds of the sweep loop alone (timed inside the library: the optically thin
    start is not part of a step)."""
    from oracle import oracle
    oracle.sphere_bgnd_solve(cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["Nmu"],
                             cfg["E_eV"], cfg["shape"], i_find_Teq=1 if cfg["find_Teq"] else 0,
                             z=cfg["z"], tol=1e-4, order=oracle.ORDER_JACOBI, flavour=flavour,
                             max_sweeps=sweeps)
    return max(oracle.last_sweep_seconds(), 1e-9)


def cpu_baseline_block(args, cfg, evals_sweep, budget_s=30.0):
    """cpu_baseline of the GPU arm's line: the full C4 sweep with all host threads, the
    thread-count sweep on a quarter-depth sample, and the faithful-structure flavour
    (per-segment tangent-shell search, six exp per evaluation: geometry.f90:125-130,
    source_background.f90:300,341) at 512 x 128 x 32."""
    from oracle import oracle
    from rabacus_b200 import configs
    nthr = host_threads()
    t_begin = time.perf_counter()
    with oracle.timing_build() as build_note:
        oracle.set_num_threads(nthr)
        t = min(cpu_sweep_seconds(cfg) for _ in range(2))
        out = {"value": evals_sweep / t, "unit": UNIT, "cores": nthr, "kind": "port",
               "sample": "best of 2 full Jacobi sweeps of the workload (%.2f s each), C "
                         "restatement of the Fortran (oracle/), OpenMP over shells with %d "
                         "threads, tangent-shell search hoisted and one exp per evaluation "
                         "(same arithmetic)" % (t, nthr),
               "build": build_note, "seconds_per_sweep": t}
        # thread-count sweep 1/2/4/../nproc on a bounded sample: the same sphere at a quarter of
        # the shells (1024 x Nnu x Nmu), which one thread finishes in ~5 s
        small = configs.c4_sphere_bgnd(max(8, args.Nl // 4), args.Nnu, args.Nmu, find_Teq=False)
        ev_small = small["nH"].size * args.Nnu * args.Nmu
        sweep, n = {}, 1
        counts = []
        while n < nthr:
            counts.append(n)
            n *= 2
        counts.append(nthr)
        for n in reversed(counts):   # most threads first: the 1-thread run is the longest
            if time.perf_counter() - t_begin > budget_s and n != nthr:
                sweep[str(n)] = None
                continue
            oracle.set_num_threads(n)
            sweep[str(n)] = ev_small / cpu_sweep_seconds(small)
        out["thread_sweep"] = {"sample": "one Jacobi sweep of the same sphere at %d shells x %d "
                                         "freqs x %d rays" % (small["nH"].size, args.Nnu, args.Nmu),
                               "evals_per_s_by_threads": sweep}
        oracle.set_num_threads(nthr)
        f = configs.c4_sphere_bgnd(512, 128, 32, find_Teq=False)
        ev_f = 512 * 128 * 32
        t_h = cpu_sweep_seconds(f, flavour=oracle.FLAVOUR_HOISTED)
        t_f = cpu_sweep_seconds(f, flavour=oracle.FLAVOUR_FAITHFUL)
        out["faithful_structure"] = {
            "sample": "one Jacobi sweep, SphereBgnd 512 x 128 x 32, %d threads" % nthr,
            "faithful_evals_per_s": ev_f / t_f, "hoisted_evals_per_s": ev_f / t_h,
            "note": "faithful = the reference's own loop structure (O(Nl) tangent-shell search "
                    "per ray segment, attenuation recomputed in each of the six integrals); "
                    "hoisted = identical arithmetic without the redundancy, the flavour every "
                    "other CPU number here uses"}
    return out


def run_reference(args, rank, world, out):
    """CPU arm: the reference's algorithm on the host cores (oracle port, -O3 build), every
    host thread this process may use -- set explicitly, because torchrun exports
    OMP_NUM_THREADS=1."""
    if rank != 0:
        return
    nthr = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(nthr)
    from oracle import oracle
    cfg = workload(args)
    evals = args.Nl * args.Nmu * args.Nnu
    sample_cfg, sample_evals, sample_txt = cfg, evals, "a full sweep of the workload"
    with oracle.timing_build() as build_note:
        nthr = oracle.set_num_threads(nthr)
        print("[bench reference] OpenMP threads: %d (host cores visible: %d); %s"
              % (nthr, host_threads(), build_note), file=sys.stderr)
        t_first = cpu_sweep_seconds(cfg)   # warm-up, and the size check of the sample
        K, W = max(1, args.steps), max(0, args.warmup)
        if t_first * (K + W) > 240.0:      # few host cores: a bounded sample of the sweep
            from rabacus_b200 import configs
            nl = max(64, args.Nl // 4)
            sample_cfg = configs.c4_sphere_bgnd(nl, args.Nnu, args.Nmu,
                                                find_Teq=bool(args.find_Teq))
            sample_evals = nl * args.Nmu * args.Nnu
            sample_txt = ("a sweep of the same sphere at %d shells (a full sweep takes %.1f s "
                          "here)" % (nl, t_first))
        for _ in range(max(0, W - 1)):
            cpu_sweep_seconds(sample_cfg)
        ts = [cpu_sweep_seconds(sample_cfg) for _ in range(K)]
    total = sum(ts)
    value = sample_evals * len(ts) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": 1e3 * total / len(ts) * (evals / sample_evals),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": bench_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthr, "kind": "port",
                         "sample": "%d step(s), each %s; OpenMP threads = %d (set explicitly, "
                                   "OMP_NUM_THREADS of the launcher ignored); C restatement of "
                                   "the Fortran with the tangent-shell search hoisted and one "
                                   "exp per evaluation (same arithmetic)"
                                   % (len(ts), sample_txt, nthr),
                         "build": build_note},
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


# ------------------------------------------------------------------------------ GPU legs
def c5_leg(args, rank, world, local_rank, dist, find_Teq):
    """BASELINE.json configs[4]: 8192 independent SphereBgnd solves (32 nH0 x 16 R x 16 z,
    512 shells x 128 freqs x 32 rays each), dealt over the ranks by a fixed shuffle, no
    communication.  wall = staging (inputs built + uploaded) + solve + read-back, max over
    ranks.  Plan creation (device buffers and the pinned host buffer the results are read
    back into, like the one-shot entries' plan cache) is reported beside it."""
    import numpy as np
    import torch
    from rabacus_b200 import configs, sharding
    from rabacus_b200.plan import Plan
    n_all = args.c5_problems
    n_nH = max(1, n_all // 256)
    idx = sharding.deal_problems(n_all, rank, world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plan = Plan(configs.GEOM_IDS["sphere_bgnd"], len(idx), 512, 128, 32, find_Teq=find_Teq,
                fixed_fcA=1.0, tol=1e-4, device=local_rank)
    plan.reserve_results()   # pinned host buffer for the batch's results, with the plan
    t1 = time.perf_counter()
    # inputs: produced on the device from (nH0, R, z) per problem, or
    # --c5-staging host: built with numpy on the host and copied
    stager = configs.c5_stage_device if args.c5_staging == "device" else configs.c5_stage
    stage = stager(plan, idx, n_nH=n_nH, find_Teq=find_Teq)
    plan.upload()
    plan.sync()
    t2 = time.perf_counter()
    info = plan.solve()
    t3 = time.perf_counter()
    res = plan.get_all(pinned=True)
    t4 = time.perf_counter()
    it_max = float(res["iters"].max())
    plan.close()
    loc = np.array([t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t1, it_max,
                    float(info["ms_columns"]), float(info["ms_nu"]), float(info["ms_cell"])])
    evals = float(info["evals"])
    if dist is not None:
        tt = torch.tensor(loc, dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        loc = tt.cpu().numpy()
        te = torch.tensor([evals], dtype=torch.float64, device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.SUM)
        evals = float(te.item())
    return {"problems": n_all, "shape": "512 shells x 128 freqs x 32 rays", "find_Teq": find_Teq,
            "wall_s": float(loc[4]), "staging_s": float(loc[1]), "solve_s": float(loc[2]),
            "readback_s": float(loc[3]), "plan_create_s": float(loc[0]),
            "evals": evals, "evals_per_s": evals / float(loc[4]), "sweeps_max": int(loc[5]),
            "staging": stage, "scaling": "strong (problems dealt over the ranks, no exchange)",
            "first_sweep_stage_ms": {"columns": float(loc[6]), "nu": float(loc[7]),
                                     "cell": float(loc[8])}}


def zones_leg():
    """Single-zone batch API (ion_solver.py:330-396,574-650) at 1e6 and 1e7
    zones through the host-pointer C ABI: wall time of the call and device time of its
    kernel (scalar per-zone solvers)."""
    import ctypes as C
    import numpy as np
    from rabacus_b200 import _lib
    L = _lib.lib()
    DP = C.POINTER(C.c_double)

    def P(a):
        return a.ctypes.data_as(DP)

    out = {}
    for n in (10 ** 6, 10 ** 7):
        u = (np.arange(n, dtype=np.float64) + 0.5) / n
        nH = 10.0 ** (-4.0 + 6.0 * u)
        nHe = nH * 0.08
        T = 10.0 ** (3.9 + 1.1 * ((u * 7919.0) % 1.0))
        H1i = 10.0 ** (-16.0 + 4.0 * ((u * 104729.0) % 1.0))
        He1i, He2i = 0.6 * H1i, 0.02 * H1i
        H1h, He1h, He2h = H1i * 5.0e-12, He1i * 8.0e-12, He2i * 2.0e-11
        one = np.ones(n)
        k = [np.empty(n) for _ in range(6)]
        _lib.check(L.rb_get_kchem(P(T), P(one), P(one), P(one), C.c_int(1), C.c_int(n),
                                  *[P(a) for a in k]))
        x = [np.empty(n) for _ in range(6)]
        t0 = time.perf_counter()
        _lib.check(L.rb_solve_pce(P(nH), P(nHe), *[P(a) for a in k], P(H1i), P(He1i), P(He2i),
                                  C.c_double(1e-6), C.c_int(n), C.c_int(0),
                                  *[P(a) for a in x[:5]]))
        t_pce = time.perf_counter() - t0
        ms_pce = L.rbi_last_zone_kernel_ms()
        t0 = time.perf_counter()
        _lib.check(L.rb_solve_pcte(P(nH), P(nHe), P(H1i), P(He1i), P(He2i), P(H1h), P(He1h),
                                   P(He2h), C.c_double(3.0), C.c_double(0.0), P(one), P(one),
                                   P(one), C.c_int(1), C.c_double(1e-6), C.c_int(n), C.c_int(0),
                                   *[P(a) for a in x]))
        t_pcte = time.perf_counter() - t0
        ms_pcte = L.rbi_last_zone_kernel_ms()
        out["%.0e" % n] = {
            "solve_pce": {"wall_s": t_pce, "kernel_ms": ms_pce, "zones_per_s_kernel":
                          n / (ms_pce * 1e-3), "zones_per_s_wall": n / t_pce},
            "solve_pcte": {"wall_s": t_pcte, "kernel_ms": ms_pcte, "zones_per_s_kernel":
                           n / (ms_pcte * 1e-3), "zones_per_s_wall": n / t_pcte,
                           "T_range": [float(x[5].min()), float(x[5].max())]}}
    out["note"] = ("rb_solve_pce / rb_solve_pcte (tol 1e-6), host pointers: the wall time includes "
                   "11 H2D and 5-6 D2H copies of pageable arrays; kernel = k_zone_pce / "
                   "k_zone_pcte, one thread per zone")
    return out


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
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the c4_find_Teq / c5 / zones legs (quick runs, profiling)")
    ap.add_argument("--c5-problems", type=int, default=8192)
    ap.add_argument("--c5-staging", default="device", choices=["device", "host"],
                    help="C5 inputs: generated on the GPU from (nH0, R, z), or numpy + H2D copy")
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
    from rabacus_b200.plan import FIELDS, Plan

    if _lib.device_count() < 1:
        raise SystemExit("bench.py --impl cuda needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W = max(3, args.warmup)
    K = max(1, args.steps)
    flush_bytes = 256 << 20
    sampler = ClockSampler(local_rank) if rank == 0 else None

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_steps(find_Teq):
        """W warm-up + K timed sweeps of C4 (cell-sharded at N > 1); returns
        (ms of the K steps, stage-time dict per step, launches in the K steps)."""
        cfg = workload(args, find_Teq)
        plan = Plan(configs.GEOM_IDS["sphere_bgnd"], 1, args.Nl, args.Nnu, args.Nmu,
                    find_Teq=find_Teq, fixed_fcA=1.0, tol=1e-4, device=local_rank,
                    max_sweeps=-(10 ** 9))  # fixed-work mode: keep sweeping past convergence
        plan.set_problem(0, cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"],
                         cfg["shape"], cfg["z"], 0.0)
        plan.upload()
        stage_src = "the K timed steps"
        if world == 1 or args.exchange == "peer":
            if world > 1:
                plan.connect_peers(rank, world)
            plan.thin()
            plan.bench_steps(1, flush_bytes)          # first sweep: work lists, lazy loading
            iw = plan.last_info()
            plan.bench_steps(W, flush_bytes)          # direct launches: stage events valid
            i0 = plan.last_info()
            if world > 1:
                # N > 1: a sharded sweep is ~0.3 ms of kernels behind seven launches ->
                # replayed as CUDA graphs (what rb_plan_solve_sharded does), one graph per
                # in-flight slot: all eight are captured here, outside the timed steps.
                # Stage times come from the W warm-up steps launched kernel by kernel.
                plan.bench_steps(8, flush_bytes, graphs=True)
            barrier()
            if sampler and not sampler.is_alive() and not sampler.stop_flag:
                sampler.start()
            i0g = plan.last_info()
            ms = plan.bench_steps(K, flush_bytes, graphs=world > 1)
            barrier()
            i1 = plan.last_info()
            if world > 1:
                st = {k: (i0[k] - iw[k]) / W for k in ("ms_columns", "ms_nu", "ms_cell",
                                                       "ms_finish")}
                stage_src = "the %d warm-up steps (direct launches; timed steps are graphs)" % W
                i0 = i0g
            else:
                st = {k: (i1[k] - i0[k]) / K for k in ("ms_columns", "ms_nu", "ms_cell",
                                                       "ms_finish")}
            launches = i1["n_launches"] - i0["n_launches"]
        else:
            send, recv = plan.shard_buffers(world)
            ext = torch.cuda.ExternalStream(plan.stream_ptr)
            flush = torch.empty(flush_bytes, dtype=torch.uint8, device="cuda")
            with torch.cuda.stream(ext):
                plan.thin()
                for _ in range(W):
                    plan.sharded_sweep(rank, world, send, recv)
                i0 = plan.last_info()
                barrier()
                if sampler and not sampler.is_alive() and not sampler.stop_flag:
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
            i1 = plan.last_info()
            st = {k: (i1[k] - i0[k]) / K for k in ("ms_columns", "ms_nu", "ms_cell", "ms_finish")}
            launches = i1["n_launches"] - i0["n_launches"]
        plan.close()
        st["source"] = stage_src
        return max_over_ranks(ms), st, launches, cfg

    evals_sweep = args.Nl * args.Nmu * args.Nnu
    ms, stages, launches, cfg = timed_steps(bool(args.find_Teq))
    value = evals_sweep * K / (ms * 1e-3)

    # ---------------------------------------------------------------- e2e
    e2e = None
    solve_ms = iters = None
    parity = None

    def one_shot(find_Teq):
        T = cfg["T"].copy()
        o = rabacus_fc.sphere_bgnd.sphere_bgnd_solve(
            cfg["Edges"].copy(), cfg["nH"].copy(), cfg["nHe"].copy(), T, args.Nmu,
            cfg["E_eV"], cfg["shape"], 1, 1.0, 1, 1, 1 if find_Teq else 0, 0, 1.0, 1.0, 1.0,
            cfg["z"], 0.0, 1e-4, args.Nl, args.Nnu)
        return o, T, dict(rabacus_fc.last_info)

    def sharded_solver(find_Teq):
        p2 = Plan(configs.GEOM_IDS["sphere_bgnd"], 1, args.Nl, args.Nnu, args.Nmu,
                  find_Teq=find_Teq, fixed_fcA=1.0, tol=1e-4, device=local_rank)
        if args.exchange == "peer":
            p2.connect_peers(rank, world)

        def call():
            p2.set_problem(0, cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"],
                           cfg["shape"], cfg["z"], 0.0)
            p2.upload()
            n = p2.solve_cell_sharded(rank, world)
            res = p2.get_problem(0)
            inf = p2.last_info()
            inf["iters"] = n
            return res, inf
        return p2, call

    def timed_solves(call, reps):
        call()  # warm (module load, allocations, plan cache)
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = call()
        barrier()
        return max_over_ranks((time.perf_counter() - t0) / reps), r

    if not args.no_e2e:
        reps = max(1, min(K, 3))
        teq = bool(args.find_Teq)
        if world == 1:
            dt, (_, _, inf) = timed_solves(lambda: one_shot(teq), reps)
        else:
            # the plan (buffers + peer mappings) is set up once, like the plan cache of
            # the one-shot entry point; every call stages, uploads, solves, downloads
            p2, call = sharded_solver(teq)
            dt, (res_sh, inf) = timed_solves(call, reps)
            # multi-GPU correctness under the driver: rank 0 solves the same problem alone
            # through the one-shot entry point and compares all twelve result fields
            if rank == 0:
                o, T1, inf1 = one_shot(teq)
                alone = dict(zip(("xH1", "xH2", "xHe1", "xHe2", "xHe3", "H1i_src", "He1i_src",
                                  "He2i_src"), o[:8]))
                alone.update(dict(zip(("H1h_src", "He1h_src", "He2h_src"), o[11:14])))
                alone["T"] = T1
                worst, same = 0.0, True
                for k in FIELDS:
                    a, b = np.asarray(res_sh[k]), np.asarray(alone[k])
                    same = same and bool(np.array_equal(a, b))
                    with np.errstate(all="ignore"):
                        r = np.where(a == b, 0.0, np.abs(a - b) / np.abs(b))
                    worst = max(worst, float(np.nanmax(r)))
                parity = {"bit_identical": same, "max_rel": worst, "fields": len(FIELDS),
                          "sweeps_sharded": int(inf["iters"]), "sweeps_n1": int(inf1["iters"]),
                          "what": "cell-sharded solve on %d GPUs vs the same solve on rank 0 "
                                  "alone (rb_sphere_bgnd_solve), all cells" % world}
            barrier()
            p2.close()
        iters = int(inf["iters"])
        solve_ms = dt * 1e3
        h2d = 8 * (args.Nl + 1 + 3 * args.Nl + args.Nnu * 9 + 9 + 2)
        d2h = 8 * 12 * args.Nl
        e2e = {"value": evals_sweep * iters / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "step": "one rb_sphere_bgnd_solve call (upload, thin "
               "start, %d sweeps to tol=1e-4, download)" % iters, "solve_wall_ms": solve_ms,
               "sweeps": iters}

    # ---------------------------------------------------------------- extras
    c4_teq = c5 = c5_teq = zones = None
    if not args.no_extras and not args.find_Teq:
        # C4's evolving-temperature flavour ("both fixed T and find_Teq")
        ms_t, st_t, _, _ = timed_steps(True)
        c4_teq = {"ms_per_step": ms_t / K, "value": evals_sweep * K / (ms_t * 1e-3), "unit": UNIT,
                  "steps": K, "stages_ms": {k: (round(v, 4) if isinstance(v, float) else v)
                                            for k, v in st_t.items()}}
        if not args.no_e2e:
            if world == 1:
                dt_t, (_, _, inf_t) = timed_solves(lambda: one_shot(True), 1)
            else:
                p3, call3 = sharded_solver(True)
                dt_t, (_, inf_t) = timed_solves(call3, 1)
                p3.close()
            c4_teq["solve"] = {"sweeps_to_converge": int(inf_t["iters"]), "wall_ms": dt_t * 1e3}
        barrier()
        c5 = c5_leg(args, rank, world, local_rank, dist, False)
        barrier()
        c5_teq = c5_leg(args, rank, world, local_rank, dist, True)
        barrier()
        if world == 1:
            zones = zones_leg()

    if sampler:  # sampled over the kernel-timed region and the solves that followed
        sampler.stop_flag = True
        if sampler.is_alive():
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
        _lib.check(_lib.lib().rbi_microbench(kind, local_rank, C.byref(v)))
        peak[name] = v.value
    props = torch.cuda.get_device_properties(local_rank)
    n_sm = props.multi_processor_count
    clk = sampler.summary() if sampler else {}
    sm_mhz_max = clk.get("sm_max_mhz") or 1965
    peak_arch_tflops = 2.0 * 64.0 * n_sm * sm_mhz_max * 1e6 / 1e12   # 64 DFMA / clk / SM
    n_seg = segments_per_sweep(cfg, args.Nmu)
    n_seg_local, ev_local = n_seg / world, evals_sweep / world
    peak_tflops = 2.0 * peak["dfma"] / 1e12
    kern = {
        "k_nu_rates": {"ms": stages["ms_nu"], "alg_ops": ev_local * OPS_PER_EVAL},
        "k_sphere_columns": {"ms": stages["ms_columns"], "alg_ops": n_seg_local * OPS_PER_SEGMENT},
    }
    for k in kern.values():
        k["tflops"] = 2.0 * k["alg_ops"] / (k["ms"] * 1e-3) / 1e12 if k["ms"] > 0 else 0.0
        k["frac"] = k["tflops"] / peak_tflops if peak_tflops > 0 else 0.0
        k["frac_arch"] = k["tflops"] / peak_arch_tflops
    # the column kernel takes the far part of every ray from a moment table: what it executes
    walked, far_evals = column_work_executed(cfg, args.Nmu)
    exec_ops_cols = (walked * OPS_PER_SEGMENT + far_evals * 3 * FAR_NT) / world
    kern["k_sphere_columns"]["exec_ops"] = exec_ops_cols
    kern["k_nu_rates"]["exec_ops"] = kern["k_nu_rates"]["alg_ops"]
    for k in kern.values():
        k["frac_executed"] = (2.0 * k["exec_ops"] / (k["ms"] * 1e-3) / 1e12 / peak_tflops
                              if k["ms"] > 0 and peak_tflops > 0 else 0.0)
    dom = max(kern, key=lambda n: kern[n]["ms"])
    traffic = traffic_pipe = None  # DRAM bytes per launch of that kernel from the committed ncu captures
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
        if world == 1:
            traffic = tj["bytes_per_launch"].get(dom)
            traffic_pipe = tj.get("bytes_per_launch_in_pipeline", {}).get(dom)
    except Exception:
        traffic = traffic_pipe = None
    # SURVEY 8(d)'s formula with the LIBRARY rates measured in this run:
    #   T_roof = N_exp/R_exp + N_sqrt/R_sqrt + (N_fma,nu + N_fma,ray)/R_fma + N_cell c_cell/R_fma
    c_cell = 1.5e5 if args.find_Teq else 1.0e3
    n_exp, n_sqrt = ev_local, n_seg_local
    n_fma = 9.0 * ev_local + 6.0 * (2.0 * n_seg_local)   # both rays of a pair visit each segment
    t_roof = (n_exp / peak["exp"] + n_sqrt / peak["sqrt"] + n_fma / peak["dfma"]
              + (args.Nl / world) * c_cell / peak["dfma"])
    step_ops = ev_local * OPS_PER_EVAL + n_seg_local * OPS_PER_SEGMENT
    step_ops_exec = ev_local * OPS_PER_EVAL + exec_ops_cols
    roofline = {
        "bound": "fp64", "kernel": dom, "achieved": kern[dom]["tflops"], "peak": peak_tflops,
        "unit": "TFLOP/s", "frac": kern[dom]["frac"],
        "peak_architectural": peak_arch_tflops, "frac_architectural": kern[dom]["frac_arch"],
        "frac_survey_model": t_roof / (ms / K * 1e-3),
        "survey_model": {"T_roof_ms": t_roof * 1e3, "T_measured_ms": ms / K, "N_exp": n_exp,
                         "N_sqrt": n_sqrt, "N_fma": n_fma, "c_cell": c_cell,
                         "R_exp": peak["exp"], "R_sqrt": peak["sqrt"], "R_fma": peak["dfma"]},
        "note": "frac = this implementation's own FP64 instruction recipe (9 per shell edge, 11 "
                "per evaluation, x2 flop) / the kernel's launch time, against the DFMA rate "
                "measured in this run; frac_architectural uses 64 DFMA/clk/SM x SMs x max clock. "
                "frac_survey_model = T_roof / T_measured with SURVEY 8(d)'s formula and the "
                "measured rates of CUDA's exp()/sqrt() LIBRARY functions: it exceeds 1 because "
                "the kernels do not call those functions (sqrt is a MUFU seed + one correction, "
                "5 FP64 instead of ~14; exp a 1024-entry table + cubic, 7 instead of ~20), so "
                "the library rates are not bounds for this implementation -- the DFMA issue "
                "rate is.  k_sphere_columns: its `frac` counts the reference algorithm's work "
                "(every shell edge above a ray's tangent point) and exceeds 1 because the kernel "
                "takes the far part of every ray from a per-sweep moment table (DESIGN.md 4); "
                "`frac_executed` counts the edges it walks and its series evaluations",
        "traffic": traffic, "traffic_in_pipeline": traffic_pipe,
        "traffic_source": "profiles/ncu_traffic.json: ncu --set full, dram__bytes_read.sum + "
                          "dram__bytes_write.sum per launch.  `traffic` is ncu's default (L2 "
                          "flushed before every replay: the kernel re-reads the 33.5 MB ray-record "
                          "buffer from DRAM); `traffic_in_pipeline` is the same capture with "
                          "--cache-control none, i.e. what the sweep sees: the buffer the column "
                          "kernel just wrote is read back from L2.  Algorithmic bytes per launch: "
                          "33.5 MB of ray records + < 1 MB of state",
        "peak_source": "measured in this run: DFMA micro-benchmark (rbi_microbench: 64 FMAs per "
                       "loop trip on 32 chains, >= 50 ms), MEASURED_PEAKS.json has no FP64 figure",
        "ms_per_launch": kern[dom]["ms"], "stage_times_from": stages["source"],
        "kernels": {n: {"ms": round(v["ms"], 4), "tflops": round(v["tflops"], 3),
                        "frac": round(v["frac"], 4), "frac_arch": round(v["frac_arch"], 4),
                        "frac_executed": round(v["frac_executed"], 4)}
                    for n, v in kern.items()},
        "ms_cell_solve": round(stages["ms_cell"], 4), "ms_finish": round(stages["ms_finish"], 4),
        "step_frac_of_fp64_peak": (2.0 * step_ops / (ms / K * 1e-3) / 1e12) / peak_tflops,
        "step_frac_architectural": (2.0 * step_ops / (ms / K * 1e-3) / 1e12) / peak_arch_tflops,
        "step_frac_executed": (2.0 * step_ops_exec / (ms / K * 1e-3) / 1e12) / peak_tflops,
        "columns_executed": {"edges_walked": walked, "far_series_evaluations": far_evals,
                             "edges_of_the_reference_algorithm": n_seg},
        "microbench_ops_per_s": peak, "segments_per_sweep": n_seg,
    }

    # ---------------------------------------------------------------- CPU baseline
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            os.environ["OMP_NUM_THREADS"] = str(host_threads())
            cpu = cpu_baseline_block(args, cfg, evals_sweep)
        except Exception as e:  # the CPU leg must never break the GPU line
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                   "sample": "failed: %r" % (e,)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args),
        "parallelism": ("all cells on one GPU" if world == 1 else
                        ("cells dealt over %d GPUs in 32-cell blocks, fractions exchanged by remote "
                         "stores over NVLink peer memory + flag barrier (no collective per sweep), "
                         "sweeps replayed as CUDA graphs" % world
                         if args.exchange == "peer" else
                         "cells dealt cyclically over %d GPUs, NCCL all-gather of 18 state fields "
                         "per sweep" % world)),
        "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        "clocks": sampler.summary() if sampler else None,
        "solve": {"sweeps_to_converge": iters, "wall_ms": solve_ms},
        "c4_find_Teq": c4_teq, "c5": c5, "c5_find_Teq": c5_teq, "zones": zones,
        "parity_vs_n1": parity,
    }
    print(json.dumps(line), file=out, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
