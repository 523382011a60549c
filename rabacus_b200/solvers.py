"""Public solver classes -- the host-side mirror of the reference's wrappers.

Same constructors, keyword arguments, iteration control and result attributes
as ``ra.SphereBgnd`` (``rabacus/f2py/sphere_bgnd.py:13-277``), ``ra.SphereStromgren``
(``sphere_stromgren.py:13-277``), ``ra.Slab2Bgnd`` (``slab_bgnd.py:17-260``),
``ra.SlabPln`` / ``ra.Slab2Pln`` (``slab_plane.py:17-484``) and the single-zone
``Solve_CE/PCE/PCTE`` (``ion_solver.py``), ``ChemistryRates`` / ``CoolingRates``
(``chem_cool_rates.py``).  Quantities are plain float64 ndarrays in cgs
(``quantities`` is not a dependency): lengths cm, densities cm^-3, T in K,
rates 1/s and erg/s.  The numerics are the CUDA C ABI behind
:mod:`rabacus_b200.rabacus_fc`.
"""
import numpy as np

from . import post, rabacus_fc

_REC = {"fixed": 1, "outward": 2, "radial": 3, "isotropic": 4, "ray": 2}


def _arr(a):
    return np.array(a, dtype=np.float64, copy=True).ravel()


class _GeomBase(object):
    """Shared input handling (``sphere_base.py:18-110``, ``slab_base.py:32-119``)."""

    _rec_allowed = ("fixed", "outward", "radial", "isotropic")
    _src_type = None

    def _setup(self, Edges, T, nH, nHe, rad_src, rec_meth, fixed_fcA, atomic_fit_name,
               find_Teq, z, Hz, verbose, tol, thin):
        if rad_src.source_type != self._src_type:
            raise ValueError("source type needs to be %s" % self._src_type)
        self.Edges, self.T, self.nH, self.nHe = _arr(Edges), _arr(T), _arr(nH), _arr(nHe)
        self.rad_src = rad_src
        self.rec_meth, self.fixed_fcA = rec_meth, float(fixed_fcA)
        self.atomic_fit_name = atomic_fit_name
        self.find_Teq, self.verbose, self.tol, self.thin = find_Teq, verbose, tol, thin
        self.z = z if find_Teq else 0.0
        assert self.Edges.size == self.T.size + 1
        assert self.T.shape == self.nH.shape == self.nHe.shape
        if rec_meth not in self._rec_allowed:
            raise ValueError("rec_meth must be one of %r" % (self._rec_allowed,))
        if find_Teq and z is None:
            raise ValueError("need to supply redshift if finding equilibrium temperature")
        self.Hz = 0.0 if Hz is None else float(Hz)
        # format_for_fortran
        self.E_eV = np.ascontiguousarray(rad_src.E, dtype=np.float64)
        self.shape = np.ascontiguousarray(rad_src.shape, dtype=np.float64)
        self.i_rec_meth = _REC[rec_meth]
        if rad_src.px_fit_type != "verner":
            raise ValueError("px_fit_type must be 'verner'")
        self.i_photo_fit = 1
        if atomic_fit_name != "hg97":
            raise ValueError("atomic_fit_name must be 'hg97'")
        self.i_rate_fit = 1
        self.i_find_Teq = 1 if find_Teq else 0
        self.i_thin = 1 if thin else 0
        self.Nl, self.Nnu = self.nH.size, self.E_eV.size

    def _attach(self, out):
        (self.xH1, self.xH2, self.xHe1, self.xHe2, self.xHe3,
         self.H1i_src, self.He1i_src, self.He2i_src,
         self.H1i_rec, self.He1i_rec, self.He2i_rec,
         self.H1h_src, self.He1h_src, self.He2h_src,
         self.H1h_rec, self.He1h_rec, self.He2h_rec) = out
        self.H1i = self.H1i_src + self.H1i_rec
        self.He1i = self.He1i_src + self.He1i_rec
        self.He2i = self.He2i_src + self.He2i_rec
        self.H1h = self.H1h_src + self.H1h_rec
        self.He1h = self.He1h_src + self.He1h_rec
        self.He2h = self.He2h_src + self.He2h_rec
        info = dict(rabacus_fc.last_info)
        self.itr = info.get("iters")  # documented at sphere_bgnd.py:118, never set there
        self.info = info

    def _post_common(self, dl):
        self.ne = self.nH * self.xH2 + self.nHe * (self.xHe2 + 2.0 * self.xHe3)
        kA = ChemistryRates(self.T, 1.0, 1.0, 1.0, fit_name=self.atomic_fit_name)
        with np.errstate(divide="ignore", invalid="ignore"):
            self.H1i_eff = kA.reH2 * self.ne * self.xH2 / self.xH1 - kA.ciH1 * self.ne
            self.He1i_eff = kA.reHe2 * self.ne * self.xHe2 / self.xHe1 - kA.ciHe1 * self.ne
            self.He2i_eff = kA.reHe3 * self.ne * self.xHe3 / self.xHe2 - kA.ciHe2 * self.ne
        self.dNH1 = dl * self.nH * self.xH1
        self.dNHe1 = dl * self.nHe * self.xHe1
        self.dNHe2 = dl * self.nHe * self.xHe2
        sig_th = [float(rabacus_fc.return_sigma(X, [rabacus_fc.return_E_th(X)])[0])
                  for X in range(3)]
        self.dtau_H1_th = self.dNH1 * sig_th[0]
        self.dtau_He1_th = self.dNHe1 * sig_th[1]
        self.dtau_He2_th = self.dNHe2 * sig_th[2]


class _SphereBase(_GeomBase):
    def _post(self):
        """``sphere_base.py:115-206``"""
        self.dr = np.abs(self.Edges[1:] - self.Edges[:-1])
        self.r_c = (self.Edges[:-1] + self.Edges[1:]) * 0.5
        self._post_common(self.dr)
        self.NH1_thru = np.sum(self.dr * self.nH * self.xH1) * 2
        self.NHe1_thru = np.sum(self.dr * self.nHe * self.xHe1) * 2
        self.NHe2_thru = np.sum(self.dr * self.nHe * self.xHe2) * 2
        with np.errstate(divide="ignore"):
            self.logNH1_thru = np.log10(self.NH1_thru)
            self.logNHe1_thru = np.log10(self.NHe1_thru)
            self.logNHe2_thru = np.log10(self.NHe2_thru)
        (self.NH_b, self.NHe_b, self.NH1_b, self.NHe1_b, self.NHe2_b) = \
            rabacus_fc.sphere_base.set_column_vs_impact(
                self.xH1, self.xH2, self.xHe1, self.xHe2, self.xHe3, self.Edges, self.nH,
                self.nHe, self.T, self.Nl)
        self.NH2_b = self.NH_b - self.NH1_b
        self.NHe3_b = self.NHe_b - self.NHe1_b - self.NHe2_b


class SphereBgnd(_SphereBase):
    """Sphere in a uniform background (``rabacus/f2py/sphere_bgnd.py:13-277``)."""

    _src_type = "background"

    def __init__(self, Edges, T, nH, nHe, rad_src, rec_meth="fixed", fixed_fcA=1.0,
                 atomic_fit_name="hg97", find_Teq=False, z=None, Hz=None, em_H1_fac=1.0,
                 em_He1_fac=1.0, em_He2_fac=1.0, verbose=False, tol=1.0e-4, thin=False,
                 Nmu=32, sweep_order=None):
        """``sweep_order`` (not in the reference): ``None`` / ``"jacobi"`` -- every cell sees the
        pre-sweep state --, or ``"gs"`` -- the reference's own sequential order
        (``sphere_bgnd.f90:753-913`` with one OpenMP thread; slow on a GPU)."""
        self.Nmu = Nmu
        self.sweep_order = sweep_order
        self.em_H1_fac, self.em_He1_fac, self.em_He2_fac = em_H1_fac, em_He1_fac, em_He2_fac
        self._setup(Edges, T, nH, nHe, rad_src, rec_meth, fixed_fcA, atomic_fit_name,
                    find_Teq, z, Hz, verbose, tol, thin)
        self.call_fortran_solver()
        self._post()

    def call_fortran_solver(self):
        import os
        old = os.environ.get("RB_SPHERE_BGND_ORDER")
        if self.sweep_order is not None:
            os.environ["RB_SPHERE_BGND_ORDER"] = "gs" if str(self.sweep_order).lower().startswith("g") else "jacobi"
        try:
            self._call()
        finally:
            if self.sweep_order is not None:
                if old is None:
                    os.environ.pop("RB_SPHERE_BGND_ORDER", None)
                else:
                    os.environ["RB_SPHERE_BGND_ORDER"] = old

    def _call(self):
        out = rabacus_fc.sphere_bgnd.sphere_bgnd_solve(
            self.Edges, self.nH, self.nHe, self.T, self.Nmu, self.E_eV, self.shape,
            self.i_rec_meth, self.fixed_fcA, self.i_photo_fit, self.i_rate_fit,
            self.i_find_Teq, self.i_thin, self.em_H1_fac, self.em_He1_fac, self.em_He2_fac,
            self.z, self.Hz, self.tol, self.Nl, self.Nnu)
        self._attach(out)


class SphereStromgren(_SphereBase):
    """Sphere with a central point source (``rabacus/f2py/sphere_stromgren.py:13-277``)."""

    _src_type = "point"

    def __init__(self, Edges, T, nH, nHe, rad_src, rec_meth="fixed", fixed_fcA=1.0,
                 atomic_fit_name="hg97", find_Teq=False, z=None, Hz=None, em_H1_fac=1.0,
                 em_He1_fac=1.0, em_He2_fac=1.0, verbose=False, tol=1.0e-4, thin=False,
                 Nmu=32):
        self.Nmu = Nmu
        self.em_H1_fac, self.em_He1_fac, self.em_He2_fac = em_H1_fac, em_He1_fac, em_He2_fac
        self._setup(Edges, T, nH, nHe, rad_src, rec_meth, fixed_fcA, atomic_fit_name,
                    find_Teq, z, Hz, verbose, tol, thin)
        self.call_fortran_solver()
        self._post()

    def call_fortran_solver(self):
        out = rabacus_fc.sphere_stromgren.sphere_stromgren_solve(
            self.Edges, self.nH, self.nHe, self.T, self.Nmu, self.E_eV, self.shape,
            self.i_rec_meth, self.fixed_fcA, self.i_photo_fit, self.i_rate_fit,
            self.i_find_Teq, self.i_thin, self.em_H1_fac, self.em_He1_fac, self.em_He2_fac,
            self.z, self.Hz, self.tol, self.Nl, self.Nnu)
        self._attach(out)


class _SlabBase(_GeomBase):
    _rec_allowed = ("fixed", "ray")

    def _post(self):
        """``slab_base.py:286-437`` (geometry and column bookkeeping)"""
        self.dl = self.Edges[1:] - self.Edges[:-1]
        self.z_c = self.Edges[:-1] + 0.5 * self.dl
        self._post_common(self.dl)
        self.NH1_thru = np.sum(self.dNH1)
        self.NHe1_thru = np.sum(self.dNHe1)
        self.NHe2_thru = np.sum(self.dNHe2)
        with np.errstate(divide="ignore"):
            self.logNH1_thru = np.log10(self.NH1_thru)
            self.logNHe1_thru = np.log10(self.NHe1_thru)
            self.logNHe2_thru = np.log10(self.NHe2_thru)
        self.NH1_c = np.cumsum(self.dNH1) - 0.5 * self.dNH1
        self.NHe1_c = np.cumsum(self.dNHe1) - 0.5 * self.dNHe1
        self.NHe2_c = np.cumsum(self.dNHe2) - 0.5 * self.dNHe2
        # slab_base.py:300-311, 369-431: mean molecular weight, central values, averages
        self.dz = self.dl
        self.mu = post.mean_molecular_weight(self.xH2, self.xHe2, self.xHe3)
        self.avg = post.slab_averages(self)


class Slab2Bgnd(_SlabBase):
    """Slab irradiated from both sides by an isotropic background
    (``rabacus/f2py/slab_bgnd.py:17-260``)."""

    _src_type = "background"

    def __init__(self, Edges, T, nH, nHe, rad_src, rec_meth="fixed", fixed_fcA=1.0,
                 atomic_fit_name="hg97", find_Teq=False, z=None, Hz=None, verbose=False,
                 tol=1.0e-4, thin=False):
        self._setup(Edges, T, nH, nHe, rad_src, "fixed" if rec_meth == "fixed" else rec_meth,
                    fixed_fcA, atomic_fit_name, find_Teq, z, Hz, verbose, tol, thin)
        if self.Nl % 2 != 0:  # slab_bgnd.py:180-182
            raise ValueError("the number of layers must be even")
        self.i_rec_meth = 1 if rec_meth == "fixed" else 2
        self.call_fortran_solver()
        self._post()

    def call_fortran_solver(self):
        out = rabacus_fc.slab_bgnd.slab_bgnd_solve(
            self.Edges, self.nH, self.nHe, self.T, self.E_eV, self.shape, self.i_rec_meth,
            self.fixed_fcA, self.i_photo_fit, self.i_rate_fit, self.i_find_Teq, self.i_thin,
            self.z, self.Hz, self.tol, self.Nl, self.Nnu)
        self._attach(out)


class SlabPln(_SlabBase):
    """Slab irradiated from one side by a plane source
    (``rabacus/f2py/slab_plane.py:17-243``)."""

    _src_type = "plane"
    _i_2side = 0

    def __init__(self, Edges, T, nH, nHe, rad_src, rec_meth="fixed", fixed_fcA=1.0,
                 atomic_fit_name="hg97", find_Teq=False, z=None, Hz=None, verbose=False,
                 tol=1.0e-4, thin=False):
        self._setup(Edges, T, nH, nHe, rad_src, rec_meth, fixed_fcA, atomic_fit_name,
                    find_Teq, z, Hz, verbose, tol, thin)
        self.i_rec_meth = 1 if rec_meth == "fixed" else 2
        self.call_fortran_solver()
        self._post()

    def call_fortran_solver(self):
        out = rabacus_fc.slab_plane.slab_plane_solve(
            self.Edges, self.nH, self.nHe, self.T, self.E_eV, self.shape, self._i_2side,
            self.i_rec_meth, self.fixed_fcA, self.i_photo_fit, self.i_rate_fit,
            self.i_find_Teq, self.i_thin, self.z, self.Hz, self.tol, self.Nl, self.Nnu)
        self._attach(out)


class Slab2Pln(SlabPln):
    """Slab irradiated from both sides by plane sources
    (``rabacus/f2py/slab_plane.py:246-484``)."""

    _i_2side = 1


# ------------------------------------------------------------------ single zone
class ChemistryRates(object):
    """``rabacus/f2py/chem_cool_rates.py:17-160`` (hg97 only)."""

    def __init__(self, T, fcA_H2, fcA_He2, fcA_He3, H1i=None, He1i=None, He2i=None,
                 fit_name="hg97"):
        if fit_name != "hg97":
            raise ValueError("fit_name must be 'hg97'")
        self.fit_name = fit_name
        self.T = np.atleast_1d(np.asarray(T, dtype=np.float64))
        n = self.T.size
        self.fcA_H2, self.fcA_He2, self.fcA_He3 = [
            np.ones(n) * np.asarray(f, dtype=np.float64) for f in (fcA_H2, fcA_He2, fcA_He3)]
        self.H1i = np.zeros(n) if H1i is None else np.ones(n) * H1i
        self.He1i = np.zeros(n) if He1i is None else np.ones(n) * He1i
        self.He2i = np.zeros(n) if He2i is None else np.ones(n) * He2i
        (self.reH2, self.reHe2, self.reHe3, self.ciH1, self.ciHe1, self.ciHe2) = \
            rabacus_fc.chem_cool_rates.get_kchem_v(self.T, self.fcA_H2, self.fcA_He2,
                                                   self.fcA_He3, 1, n)


class CoolingRates(object):
    """``rabacus/f2py/chem_cool_rates.py:240-400`` (hg97 only)."""

    def __init__(self, T, fcA_H2, fcA_He2, fcA_He3, H1h=None, He1h=None, He2h=None,
                 fit_name="hg97"):
        if fit_name != "hg97":
            raise ValueError("fit_name must be 'hg97'")
        self.fit_name = fit_name
        self.T = np.atleast_1d(np.asarray(T, dtype=np.float64))
        n = self.T.size
        self.fcA_H2, self.fcA_He2, self.fcA_He3 = [
            np.ones(n) * np.asarray(f, dtype=np.float64) for f in (fcA_H2, fcA_He2, fcA_He3)]
        self.H1h = np.zeros(n) if H1h is None else np.ones(n) * H1h
        self.He1h = np.zeros(n) if He1h is None else np.ones(n) * He1h
        self.He2h = np.zeros(n) if He2h is None else np.ones(n) * He2h
        (self.recH2, self.recHe2, self.recHe3, self.cicH1, self.cicHe1, self.cicHe2,
         self.cecH1, self.cecHe2, self.bremss, self.compton) = \
            rabacus_fc.chem_cool_rates.get_kcool_v(self.T, self.fcA_H2, self.fcA_He2,
                                                   self.fcA_He3, 1, n)


class Solve_CE(object):
    """Collisional equilibrium (``rabacus/f2py/ion_solver.py:95-165``)."""

    def __init__(self, kchem):
        nn = kchem.T.size
        (self.H1, self.H2, self.He1, self.He2, self.He3) = rabacus_fc.ion_solver.solve_ce_v(
            kchem.reH2, kchem.reHe2, kchem.reHe3, kchem.ciH1, kchem.ciHe1, kchem.ciHe2, nn)


class Solve_PCE(object):
    """Photo-collisional equilibrium (``rabacus/f2py/ion_solver.py:330-396``).
    As in the reference, more than one zone uses the lock-step ``_v`` solver."""

    def __init__(self, nH_in, nHe_in, kchem, tol=1.0e-10):
        nH = np.atleast_1d(np.asarray(nH_in, dtype=np.float64))
        nHe = np.atleast_1d(np.asarray(nHe_in, dtype=np.float64))
        if nH.shape != nHe.shape:
            raise ValueError("nH and nHe must have the same shape")
        nn = nH.size
        solve = rabacus_fc.ion_solver.solve_pce_v if nn > 1 else \
            (lambda *a: tuple(np.array([v]) for v in rabacus_fc.ion_solver.solve_pce_s(*a)))
        (self.H1, self.H2, self.He1, self.He2, self.He3) = solve(
            nH, nHe, kchem.reH2, kchem.reHe2, kchem.reHe3, kchem.ciH1, kchem.ciHe1,
            kchem.ciHe2, kchem.H1i, kchem.He1i, kchem.He2i, tol, nn)
        self.nH, self.nHe = nH, nHe
        self.ne = self.H2 * nH + (self.He2 + 2.0 * self.He3) * nHe


class Solve_PCTE(object):
    """Photo-collisional-thermal equilibrium (``rabacus/f2py/ion_solver.py:574-650``)."""

    def __init__(self, nH_in, nHe_in, kchem, kcool, z, Hz=None, tol=1.0e-10):
        nH = np.atleast_1d(np.asarray(nH_in, dtype=np.float64))
        nHe = np.atleast_1d(np.asarray(nHe_in, dtype=np.float64))
        if nH.shape != nHe.shape:
            raise ValueError("nH and nHe must have the same shape")
        self.Hz = 0.0 if Hz is None else float(Hz)
        nn = nH.size
        solve = rabacus_fc.ion_solver.solve_pcte_v if nn > 1 else \
            (lambda *a: tuple(np.array([v]) for v in rabacus_fc.ion_solver.solve_pcte_s(*a)))
        (self.H1, self.H2, self.He1, self.He2, self.He3, self.Teq) = solve(
            nH, nHe, kchem.H1i, kchem.He1i, kchem.He2i, kcool.H1h, kcool.He1h, kcool.He2h,
            z, self.Hz, kchem.fcA_H2, kchem.fcA_He2, kchem.fcA_He3, 1, tol, nn)
        self.nH, self.nHe = nH, nHe
        self.ne = self.H2 * nH + (self.He2 + 2.0 * self.He3) * nHe
