"""Device-resident batches of problems (``rb_plan_*`` in ``include/rabacus_b200.h``).

A :class:`Plan` holds ``n_problems`` problems of identical shape in HBM.  It is
used three ways:

* one GPU, one call: ``plan.solve()``;
* parameter sweeps: each rank builds a plan over its share of the
  problems -- no communication at all (:func:`shard_problems`);
* cell sharding: every rank holds the whole problem, sweeps the
  shells ``il = rank, rank+G, ...`` (cyclic: ray cost varies smoothly with
  ``il``, so this balances it), then the updated ionisation fractions,
  temperature and rates are all-gathered over NCCL/NVLink
  (``torch.distributed``) and every rank runs the same convergence test on the
  full arrays (:meth:`Plan.solve_cell_sharded`).  This is a Jacobi iteration;
  it reproduces the single-GPU iterates exactly.
"""
import ctypes as C

import numpy as np

from ._lib import (GEOM_SLAB_BGND, GEOM_SLAB_PLANE_2, check, lib, rb_info, rb_plan_desc)

_DP = C.POINTER(C.c_double)

FIELDS = ("xH1", "xH2", "xHe1", "xHe2", "xHe3", "T", "H1i_src", "He1i_src", "He2i_src",
          "H1h_src", "He1h_src", "He2h_src")
REC_FIELDS = ("H1i_rec", "He1i_rec", "He2i_rec", "H1h_rec", "He1h_rec", "He2h_rec")


def _p(a):
    return a.ctypes.data_as(_DP)


class Plan(object):
    def __init__(self, geometry, n_problems, Nl, Nnu, Nmu=0, find_Teq=False, fixed_fcA=1.0,
                 tol=1.0e-4, device=-1, max_sweeps=0, i_rec_meth=1, em_fac=(1.0, 1.0, 1.0)):
        self.desc = rb_plan_desc(int(geometry), int(n_problems), int(Nl), int(Nnu), int(Nmu),
                                 1 if find_Teq else 0, float(fixed_fcA), float(tol),
                                 int(device), int(max_sweeps), int(i_rec_meth),
                                 (C.c_double * 3)(*[float(v) for v in em_fac]), 1)
        self._h = C.c_void_p()
        check(lib().rb_plan_create(C.byref(self._h), C.byref(self.desc)))
        self.n_problems, self.Nl, self.Nnu, self.Nmu = n_problems, Nl, Nnu, Nmu
        self.geometry = geometry
        self.device = device

    def check_guards(self):
        """Checked build (``RB_LIBRARY=.../librabacus_b200_checked.so``): number of guard bytes
        behind the plan's device buffers that a kernel overwrote; -1 in the release build."""
        where = C.create_string_buffer(64)
        n = lib().rbi_plan_check_guards(self._h, where, C.c_size_t(64))
        return n, where.value.decode()

    def close(self):
        if self._h:
            n, where = self.check_guards()
            lib().rb_plan_destroy(self._h)
            self._h = C.c_void_p()
            if n > 0:
                raise RuntimeError("checked build: %d guard bytes overwritten behind %s" % (n, where))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ staging
    def set_problem(self, b, Edges, nH, nHe, T, E_eV, shape, z=0.0, Hz=0.0):
        arrs = [np.ascontiguousarray(a, dtype=np.float64)
                for a in (Edges, nH, nHe, T, E_eV, shape)]
        if arrs[0].size != self.Nl + 1 or arrs[1].size != self.Nl or arrs[4].size != self.Nnu:
            raise ValueError("array sizes do not match the plan")
        check(lib().rb_plan_set_problem(self._h, C.c_int(b), *[_p(a) for a in arrs],
                                        C.c_double(float(z)), C.c_double(float(Hz))))

    def set_batch(self, b0, Edges, nH, nHe, T, E_eV, shape, spec_index, z, Hz=None):
        """Stage ``n`` problems ``b0 .. b0+n-1`` in one call: ``Edges [n, Nl+1]``,
        ``nH / nHe / T [n, Nl]``, spectra ``E_eV / shape [nspec, Nnu]`` shared through
        ``spec_index [n]``, ``z [n]``, ``Hz [n]``."""
        A = [np.ascontiguousarray(a, dtype=np.float64) for a in (Edges, nH, nHe, T, E_eV, shape)]
        n = A[1].shape[0]
        if A[0].shape != (n, self.Nl + 1) or A[1].shape != (n, self.Nl) or \
                A[4].ndim != 2 or A[4].shape[1] != self.Nnu or A[5].shape != A[4].shape:
            raise ValueError("array shapes do not match the plan")
        si = np.ascontiguousarray(spec_index, dtype=np.int32)
        zz = np.ascontiguousarray(z, dtype=np.float64)
        hz = np.zeros(n) if Hz is None else np.ascontiguousarray(Hz, dtype=np.float64)
        check(lib().rb_plan_set_batch(self._h, C.c_int(int(b0)), C.c_int(n), *[_p(a) for a in A[:4]],
                                      C.c_int(A[4].shape[0]), _p(A[4]), _p(A[5]),
                                      si.ctypes.data_as(C.POINTER(C.c_int)), _p(zz), _p(hz)))

    def set_sphere_family(self, nH0, R_cm, z, Hz=None, core_div=100.0, slope=2.0,
                          nHe_over_nH=None, T0=1.0e4, q_min=1.0, q_max=400.0, hm12=None):
        """Describe EVERY problem of a SphereBgnd plan by (nH0, R, z): linspace edges, a flat
        core + power-law envelope, HM12 spectrum at z.  ``upload`` then produces all solver
        inputs on the device (``rb_plan_set_sphere_family``).  Returns the
        photon-energy grid [eV]."""
        from . import rad_src
        n = self.n_problems
        A = [np.ascontiguousarray(a, dtype=np.float64) for a in (nH0, R_cm, z)]
        if any(a.shape != (n,) for a in A):
            raise ValueError("nH0, R_cm, z must have one entry per problem of the plan")
        hz = None if Hz is None else np.ascontiguousarray(Hz, dtype=np.float64)
        if nHe_over_nH is None:
            nHe_over_nH = 0.25 * 0.24 / (1.0 - 0.24)
        tab = hm12 if hm12 is not None else rad_src.HM12Table()
        lam = np.ascontiguousarray(tab.log_lam_cm, dtype=np.float64)
        zt = np.ascontiguousarray(tab.z, dtype=np.float64)
        logI = np.ascontiguousarray(tab.logInu, dtype=np.float64)
        E = np.empty(self.Nnu)
        check(lib().rb_plan_set_sphere_family(
            self._h, _p(A[0]), _p(A[1]), _p(A[2]), _p(hz) if hz is not None else None,
            C.c_double(core_div), C.c_double(slope), C.c_double(nHe_over_nH), C.c_double(T0),
            C.c_double(q_min), C.c_double(q_max), C.c_int(lam.size), C.c_int(zt.size),
            _p(lam), _p(zt), _p(logI), _p(E)))
        return E

    def fetch_inputs(self, b):
        """Problem ``b``'s staged inputs as the kernels see them (solver orientation)."""
        out = dict(Edges=np.empty(self.Nl + 1), nH=np.empty(self.Nl), nHe=np.empty(self.Nl),
                   T=np.empty(self.Nl), spec=np.empty((self.Nnu, 9)), thin=np.empty(9))
        check(lib().rbi_plan_fetch_inputs(self._h, C.c_int(int(b)), *[_p(out[k]) for k in
                                          ("Edges", "nH", "nHe", "T", "spec", "thin")]))
        return out

    def set_sweep_order(self, order):
        """SphereBgnd: ``"jacobi"`` (default) or ``"gs"`` -- the reference's sequential
        Gauss-Seidel order (``sphere_bgnd.f90:753-913`` with one OpenMP thread)."""
        code = {"jacobi": 0, "gs": 1, "gauss-seidel": 1, 0: 0, 1: 1}[order]
        check(lib().rb_plan_set_sweep_order(self._h, C.c_int(code)))

    def upload(self):
        check(lib().rb_plan_upload(self._h))

    # ------------------------------------------------------------------ stages
    def thin(self):
        check(lib().rb_plan_thin(self._h))

    def sweep_local(self, cell_start=0, cell_stride=1):
        check(lib().rb_plan_sweep_local(self._h, C.c_int(cell_start), C.c_int(cell_stride)))

    def finish_sweep(self):
        n = C.c_int(0)
        check(lib().rb_plan_finish_sweep(self._h, C.byref(n)))
        return n.value

    def finish_sweep_async(self):
        """Enqueue the convergence test of the sweep just enqueued; does not wait."""
        check(lib().rb_plan_finish_sweep_async(self._h))

    def poll_sweep(self):
        """Wait for the oldest un-polled sweep; number of problems still active after it."""
        n = C.c_int(0)
        check(lib().rb_plan_poll_sweep(self._h, C.byref(n)))
        return n.value

    def sync(self):
        check(lib().rb_plan_sync(self._h))

    def solve(self):
        info = rb_info()
        check(lib().rb_plan_solve(self._h, C.byref(info)))
        return info.as_dict()

    def bench_steps(self, steps, flush_bytes=0, graphs=False):
        """Run ``steps`` sweeps; returns their summed device time [ms] (CUDA events).
        ``graphs``: replay the sweeps as CUDA graphs like :meth:`solve` does."""
        ms = C.c_double(0.0)
        check(lib().rbi_plan_bench_steps(self._h, C.c_int(int(steps)),
                                         C.c_size_t(int(flush_bytes)),
                                         C.c_int(1 if graphs else 0), C.byref(ms)))
        return ms.value

    def set_far_field(self, on):
        """Diagnostic A/B switch of THIS plan: SphereBgnd columns with (default) / without
        the far-field moment table."""
        check(lib().rbi_plan_set_far_field(self._h, C.c_int(1 if on else 0)))

    def set_libm_pow(self, on):
        """Diagnostic A/B switch of THIS plan: rate fits through libdevice pow()."""
        check(lib().rbi_plan_set_libm_pow(self._h, C.c_int(1 if on else 0)))

    def last_info(self):
        info = rb_info()
        check(lib().rb_plan_last_info(self._h, C.byref(info)))
        return info.as_dict()

    @property
    def evals_per_sweep(self):
        """cell*ray*frequency evaluations of one sweep of ONE problem"""
        return int(lib().rb_plan_evals_per_sweep(self._h))

    # ------------------------------------------------------------------ results
    def get_problem(self, b):
        outs = {k: np.zeros(self.Nl) for k in FIELDS}
        it = C.c_int(0)
        order = ("T", "xH1", "xH2", "xHe1", "xHe2", "xHe3", "H1i_src", "He1i_src",
                 "He2i_src", "H1h_src", "He1h_src", "He2h_src")
        check(lib().rb_plan_get_problem(self._h, C.c_int(b), *[_p(outs[k]) for k in order],
                                        C.byref(it)))
        outs["iters"] = it.value
        rec = {k: np.zeros(self.Nl) for k in REC_FIELDS}
        check(lib().rb_plan_get_rec(self._h, C.c_int(b), *[_p(rec[k]) for k in REC_FIELDS]))
        outs.update(rec)
        return outs

    def reserve_results(self):
        """Allocate the plan's pinned host result buffer now (``get_all(pinned=True)`` would do
        it on first use): 18 x n_problems x Nl doubles."""
        check(lib().rb_plan_reserve_results(self._h))

    def get_all(self, with_rec=False, pinned=False):
        """All problems' results in one device pass + copy: dict of [n_problems, Nl]
        arrays (input orientation) and ``iters`` [n_problems].  ``pinned``: the arrays are
        read-only VIEWS of a pinned host buffer owned by the plan (no allocation, copy at PCIe
        speed) -- valid until the next ``get_all`` or ``close``."""
        B, Nl = self.n_problems, self.Nl
        if pinned:
            po, pr, pi = _DP(), _DP(), C.POINTER(C.c_int)()
            check(lib().rb_plan_get_all_pinned(self._h, C.byref(po),
                                               C.byref(pr) if with_rec else None, C.byref(pi)))
            out = np.ctypeslib.as_array(po, shape=(12, B, Nl))
            res = {k: out[i] for i, k in enumerate(FIELDS)}
            if with_rec:
                rec = np.ctypeslib.as_array(pr, shape=(6, B, Nl))
                res.update({k: rec[i] for i, k in enumerate(REC_FIELDS)})
            res["iters"] = np.ctypeslib.as_array(pi, shape=(B,))
            for v in res.values():
                v.flags.writeable = False
            return res
        out = np.empty((12, B, Nl))
        rec = np.empty((6, B, Nl)) if with_rec else None
        iters = np.zeros(B, dtype=np.int32)
        check(lib().rb_plan_get_all(self._h, _p(out), _p(rec) if with_rec else None,
                                    iters.ctypes.data_as(C.POINTER(C.c_int))))
        res = {k: out[i] for i, k in enumerate(FIELDS)}
        if with_rec:
            res.update({k: rec[i] for i, k in enumerate(REC_FIELDS)})
        res["iters"] = iters
        return res

    # ------------------------------------------------------------------ torch interop
    def field_tensor(self, field):
        """A torch CUDA tensor aliasing one device field ([n_problems*Nl] float64)."""
        import torch
        idx = FIELDS.index(field) if isinstance(field, str) else int(field)
        ptr = lib().rb_plan_field_ptr(self._h, idx)
        n = self.n_problems * self.Nl
        dev = torch.cuda.current_device() if self.device < 0 else self.device

        class _Arr(object):
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False),
                                        "version": 2}
        return torch.as_tensor(_Arr(), device="cuda:%d" % dev)

    def n_sweep_cells(self):
        two = self.geometry in (GEOM_SLAB_PLANE_2, GEOM_SLAB_BGND)
        return self.Nl // 2 if two else self.Nl

    @property
    def stream_ptr(self):
        return lib().rb_plan_stream(self._h)

    def pack_cells(self, cell_start, cell_stride, dev_send_ptr):
        check(lib().rb_plan_pack_cells(self._h, C.c_int(cell_start), C.c_int(cell_stride),
                                       C.c_void_p(dev_send_ptr)))

    def unpack_cells(self, world, dev_recv_ptr):
        check(lib().rb_plan_unpack_cells(self._h, C.c_int(world), C.c_void_p(dev_recv_ptr)))

    def shard_buffers(self, world):
        """(send [18, n_local], recv [world, 18, n_local]) float64 CUDA tensors."""
        import torch
        ncell = self.n_sweep_cells()
        if ncell % world != 0:
            raise ValueError("number of swept cells (%d) must be divisible by world size %d"
                             % (ncell, world))
        nloc = ncell // world
        dev = torch.cuda.current_device() if self.device < 0 else self.device
        nf = len(FIELDS) + len(REC_FIELDS)
        send = torch.empty((nf, nloc), dtype=torch.float64, device="cuda:%d" % dev)
        recv = torch.empty((world, nf, nloc), dtype=torch.float64,
                           device="cuda:%d" % dev)
        return send, recv

    # ------------------------------------------------------------------ peer-memory sharding
    def connect_peers(self, rank, world, group=None):
        """Map every rank's staging buffer into this process (CUDA IPC over
        NVLink/NVSwitch peer access).  Collective over ``group``; afterwards
        :meth:`solve_cell_sharded` exchanges the fractions with remote stores from
        the cell-solver kernel and a flag barrier -- no collective per sweep."""
        import torch
        import torch.distributed as dist
        h = (C.c_ubyte * 64)()
        check(lib().rb_plan_peer_alloc(self._h, C.c_int(rank), C.c_int(world), h))
        mine = torch.tensor(list(bytes(h)), dtype=torch.uint8)
        dev = torch.cuda.current_device() if self.device < 0 else self.device
        if dist.get_backend(group) == "nccl":
            mine = mine.to("cuda:%d" % dev)
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine, group=group)
        blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh)
        buf = (C.c_ubyte * (64 * world)).from_buffer_copy(blob)
        check(lib().rb_plan_peer_connect(self._h, buf))
        self._peer = (rank, world)

    def sweep_sharded(self):
        check(lib().rb_plan_sweep_sharded(self._h))

    def gather_cells(self, rank, world, group=None):
        """One all-gather of the 18 state fields of every rank's cells (the twelve
        rates are not exchanged per sweep on the peer path)."""
        import torch
        import torch.distributed as dist
        send, recv = self.shard_buffers(world)
        ext = torch.cuda.ExternalStream(self.stream_ptr)
        # the deal of a peer-connected plan (cyclic cells, or 32-cell blocks: cell_of() in
        # csrc/rb_common.cuh) decides which cells this rank packs and where they land
        deal = lib().rb_plan_peer_deal(self._h) if getattr(self, "_peer", None) else world
        deal = deal if abs(deal) == world else world
        with torch.cuda.stream(ext):
            self.pack_cells(rank, deal, send.data_ptr())
            dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
            self.unpack_cells(deal, recv.data_ptr())
        self.sync()

    def enqueue_sharded_sweep(self, rank, world, send, recv, group=None):
        """Enqueue one cell-sharded Jacobi sweep: local cells -> pack -> NCCL
        all-gather -> unpack -> convergence test on the full arrays.  Must run with
        the plan's stream as torch's current stream (see :meth:`solve_cell_sharded`)
        so that kernels and the collective are ordered on the device without host
        synchronisation."""
        import torch.distributed as dist
        self.sweep_local(rank, world)
        self.pack_cells(rank, world, send.data_ptr())
        dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
        self.unpack_cells(world, recv.data_ptr())
        self.finish_sweep_async()

    def sharded_sweep(self, rank, world, send, recv, group=None):
        """One sharded sweep, waited for; returns the number of still-active problems."""
        self.enqueue_sharded_sweep(rank, world, send, recv, group)
        return self.poll_sweep()

    def solve_cell_sharded(self, rank, world, group=None, max_sweeps=100000, depth=3):
        """Jacobi solve with the sweep's cells dealt cyclically over ``world`` ranks.

        Requires ``torch.distributed`` to be initialised (NCCL on GPUs).  Every
        rank must hold the same staged problem.  Sweeps are enqueued ``depth`` ahead
        of the convergence read-back (sweeps enqueued after convergence are no-ops
        on the device and every rank takes the same decisions, so the collectives
        stay matched).  Returns the sweep count.
        """
        import torch
        if self.n_problems != 1:
            raise ValueError("cell sharding is defined for a single problem per plan")
        if getattr(self, "_peer", None) == (rank, world):
            # peer-memory path: the whole pipelined loop runs in the library
            info = rb_info()
            check(lib().rb_plan_solve_sharded(self._h, C.byref(info)))
            self.gather_cells(rank, world, group)
            return info.iters
        send, recv = self.shard_buffers(world)
        ext = torch.cuda.ExternalStream(self.stream_ptr)
        issued = polled = sweeps = 0
        stop = False
        with torch.cuda.stream(ext):
            self.thin()
            while True:
                if not stop and issued - polled < depth and issued < max_sweeps:
                    self.enqueue_sharded_sweep(rank, world, send, recv, group)
                    issued += 1
                    continue
                if polled >= issued:
                    break
                n_active = self.poll_sweep()
                polled += 1
                if not stop:
                    sweeps = polled
                if n_active == 0:
                    stop = True
        self.sync()
        return sweeps


def shard_problems(n_problems, rank, world):
    """Contiguous block partition of a problem batch (no communication)."""
    per = (n_problems + world - 1) // world
    lo = min(n_problems, rank * per)
    hi = min(n_problems, lo + per)
    return lo, hi


def deal_problems(n_problems, rank, world):
    """Shuffled deal of a problem batch (balances sweep counts across ranks); see
    :func:`rabacus_b200.sharding.deal_problems`."""
    from .sharding import deal_problems as _deal
    return _deal(n_problems, rank, world)

