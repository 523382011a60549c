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


def _p(a):
    return a.ctypes.data_as(_DP)


class Plan(object):
    def __init__(self, geometry, n_problems, Nl, Nnu, Nmu=0, find_Teq=False, fixed_fcA=1.0,
                 tol=1.0e-4, device=-1, max_sweeps=0):
        self.desc = rb_plan_desc(int(geometry), int(n_problems), int(Nl), int(Nnu), int(Nmu),
                                 1 if find_Teq else 0, float(fixed_fcA), float(tol),
                                 int(device), int(max_sweeps))
        self._h = C.c_void_p()
        check(lib().rb_plan_create(C.byref(self._h), C.byref(self.desc)))
        self.n_problems, self.Nl, self.Nnu, self.Nmu = n_problems, Nl, Nnu, Nmu
        self.geometry = geometry
        self.device = device

    def close(self):
        if self._h:
            lib().rb_plan_destroy(self._h)
            self._h = C.c_void_p()

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

    def sync(self):
        check(lib().rb_plan_sync(self._h))

    def solve(self):
        info = rb_info()
        check(lib().rb_plan_solve(self._h, C.byref(info)))
        return info.as_dict()

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
        return outs

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

    def solve_cell_sharded(self, rank, world, group=None, max_sweeps=100000):
        """Jacobi solve with the sweep's cells dealt cyclically over ``world`` ranks.

        Requires ``torch.distributed`` to be initialised (NCCL on GPUs).  Every
        rank must hold the same staged problem(s).  Returns the sweep count.
        """
        import torch
        import torch.distributed as dist
        if self.n_problems != 1:
            raise ValueError("cell sharding is defined for a single problem per plan")
        ncell = self.n_sweep_cells()
        if ncell % world != 0:
            raise ValueError("number of swept cells (%d) must be divisible by world size %d"
                             % (ncell, world))
        two = ncell != self.Nl
        fields = [self.field_tensor(f) for f in FIELDS]
        nloc = ncell // world
        self.thin()
        self.sync()
        sweeps = 0
        n_active = 1
        send = torch.empty((len(FIELDS), nloc), dtype=torch.float64, device=fields[0].device)
        recv = torch.empty((world, len(FIELDS), nloc), dtype=torch.float64,
                           device=fields[0].device)
        while n_active > 0 and sweeps < max_sweeps:
            self.sweep_local(rank, world)
            self.sync()  # plan stream -> torch stream hand-off
            for i, f in enumerate(fields):
                send[i] = f[:ncell].view(nloc, world)[:, rank]
            dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
            # recv[g, i, j] is field i of cell j*world + g
            full = recv.permute(1, 2, 0).reshape(len(FIELDS), ncell)
            for i, f in enumerate(fields):
                f[:ncell] = full[i]
                if two:  # mirrored half of the two-sided slabs
                    f[self.Nl - ncell:] = torch.flip(full[i], dims=(0,))
            torch.cuda.current_stream().synchronize()
            n_active = self.finish_sweep()
            sweeps += 1
        return sweeps


def shard_problems(n_problems, rank, world):
    """Contiguous block partition of a problem batch (no communication)."""
    per = (n_problems + world - 1) // world
    lo = min(n_problems, rank * per)
    hi = min(n_problems, lo + per)
    return lo, hi
