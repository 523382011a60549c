"""Host-side input producers: photon-energy grid, spectral shapes, normalisation.

Mirrors (in plain cgs ndarrays -- ``quantities`` is not a dependency) the parts
of ``rabacus/rad_src`` the geometric solvers consume:

* ``Source.set_photon_arrays`` -- log-spaced grid with the samples nearest the
  H I / He I / He II thresholds snapped onto them (``rad_src/source.py:178-259``);
* spectral shapes ``hm12`` (``rad_src/hm12.py`` + ``uv_bgnd/hm12.py:199-303``),
  ``thermal`` (``rad_src/thermal.py``), ``powerlaw`` (``rad_src/powerlaw.py``),
  ``monochromatic`` (``rad_src/monochromatic.py``) and ``user``;
* the source classes ``BackgroundSource`` (``rad_src/background.py:249-325``),
  ``PointSource`` (``rad_src/point.py``) and ``PlaneSource`` (``rad_src/plane.py``)
  with their optically-thin integrals and ``normalize_*`` methods.

These run once per problem on the host; they are inputs to the hot path, not
part of it (SURVEY 2b "input producer").
"""
import os

import numpy as np

# rabacus/f2py/physical_constants.f90:7-21 (the values the Fortran integrals use)
PI = 3.141592653589793
H_EV_S = 4.135667603369524e-15
H_ERG_S = 6.62606957e-27
C_CM_S = 2.99792458e10
KB_ERG_K = 1.3806488e-16
EV_2_ERG = 1.60217653e-12
ERG_2_EV = 6.241509479607719e11
KPC_CM = 3.0856775814913673e21  # stated explicitly: SURVEY 8(c), unit factors are unpinned

# Verner 96 thresholds, rabacus/atomic/verner/photox/photo.dat:1-3
ETH_H1, ETH_HE1, ETH_HE2 = 13.60, 24.59, 54.42

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def trap(x, y):
    """``rabacus/utils/utils.py:55-61`` / ``utils.f90:11-21``"""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    return 0.5 * np.sum((y[1:] + y[:-1]) * (x[1:] - x[:-1]))


def photon_energies(q_min, q_max, Nnu, segmented=True):
    """Energy samples in eV (``rad_src/source.py:178-259``)."""
    log_q = np.linspace(np.log10(q_min), np.log10(q_max), Nnu)
    q = 10.0 ** log_q
    idx = {}
    if segmented and Nnu > 1:
        q_He1, q_He2 = ETH_HE1 / ETH_H1, ETH_HE2 / ETH_H1
        if q.min() > 1.0:
            raise ValueError("q_min must be <= 1 for segmented spectra")
        if q.max() < q_He2:
            raise ValueError("q_max must be >= %g for segmented spectra" % q_He2)
        for name, qt in (("H1", 1.0), ("He1", q_He1), ("He2", q_He2)):
            i = int(np.argmin(np.abs(log_q - np.log10(qt))))
            q[i] = qt
            idx[name] = i
    return q * ETH_H1, idx


class HM12Table(object):
    """Haardt & Madau 2012 UVB table (``rabacus/uv_bgnd/hm12.py:175-303``)."""

    _cache = None

    def __init__(self):
        if HM12Table._cache is None:
            HM12Table._cache = dict(np.load(os.path.join(_DATA, "hm12_uvb.npz")))
        d = HM12Table._cache
        self.z = d["z"]
        Inu = d["Inu"]
        pos = Inu > 0
        self.logInu = np.where(pos, np.log10(np.where(pos, Inu, 1.0)),
                               np.log10(Inu[pos].min()))  # hm12.py:213-219
        self.log_lam_cm = np.log10(d["lam_A"] * 1.0e-8)

    def spectrum_E(self, z, E_eV):
        """Inu [erg/(s cm^2 Hz sr)] at photon energies ``E_eV``: linear in z,
        log-log in wavelength (``hm12.py:239-303``)."""
        z = float(z)
        i_lo = int(np.argmin(np.abs(self.z - z)))
        if self.z[i_lo] < z:
            i_hi = i_lo + 1
        else:
            i_hi = i_lo
            i_lo = i_lo - 1
        if i_lo < 0:  # z == table minimum
            i_lo, i_hi = 0, 1
        z_frac = (z - self.z[i_lo]) / (self.z[i_hi] - self.z[i_lo])
        log_Inu = self.logInu[:, i_lo] + z_frac * (self.logInu[:, i_hi] - self.logInu[:, i_lo])
        lam_cm = C_CM_S / (np.asarray(E_eV, dtype=np.float64) / H_EV_S)
        return 10.0 ** np.interp(np.log10(lam_cm), self.log_lam_cm, log_Inu)


def _sigma(species, E):
    from . import rabacus_fc
    return rabacus_fc.return_sigma(species, E)


class Source(object):
    """Common spectrum handling (``rad_src/source.py``)."""

    source_type = None
    px_fit_type = "verner"

    def _initialize(self, q_min, q_max, spectrum_type, Nnu, segmented, z, T_eff, alpha,
                    user_E, user_shape):
        self.spectrum_type = spectrum_type
        self.z, self.T_eff, self.alpha = z, T_eff, alpha
        self.monochromatic = spectrum_type == "monochromatic"
        if self.monochromatic:
            if q_min != q_max:
                raise ValueError("source type = monochromatic, but q_min != q_max")
            Nnu, segmented = 1, False
        if spectrum_type == "user":
            self.E = np.array(user_E, dtype=np.float64)
            self.th_index = {}
        else:
            self.E, self.th_index = photon_energies(q_min, q_max, Nnu, segmented)
        self.Nnu = self.E.size
        self.q = self.E / ETH_H1
        self.nu = self.E / H_EV_S
        self.lam = C_CM_S / self.nu
        if spectrum_type == "hm12":
            if z is None:
                raise ValueError("if source type = hm12, must provide z")
            yvals = HM12Table().spectrum_E(z, self.E)
        elif spectrum_type == "thermal":
            if T_eff is None:
                raise ValueError("if source type = thermal, must provide T_eff")
            t1 = 2.0 * H_ERG_S * self.nu ** 3 / C_CM_S ** 2
            yvals = t1 / (np.exp(H_ERG_S * self.nu / (KB_ERG_K * T_eff)) - 1.0)
        elif spectrum_type == "powerlaw":
            if alpha is None:
                raise ValueError("if source type = powerlaw, must provide alpha")
            yvals = (self.nu / self.nu[0]) ** alpha
        elif spectrum_type == "monochromatic":
            yvals = np.ones(1)
        elif spectrum_type == "user":
            yvals = np.array(user_shape, dtype=np.float64)
            if yvals.size != self.E.size:
                raise ValueError("E and shape must be same size for source_type = user")
        else:
            raise ValueError("spectrum_type not recognised: %r" % spectrum_type)
        self.shape = np.ones(self.Nnu) * yvals
        self.sigma_H1 = _sigma(0, self.E)
        self.sigma_He1 = _sigma(1, self.E)
        self.sigma_He2 = _sigma(2, self.E)

    # ---- optically thin integrals; `base` is the photon-number integrand per eV
    def _base(self, r=None):
        raise NotImplementedError

    def _integrate(self, y):
        return float(y[0]) if self.Nnu == 1 else trap(self.E, y)

    def thin_rates(self, r=None):
        """(H1i, He1i, He2i [1/s], H1h, He1h, He2h [erg/s]) with no attenuation."""
        b = self._base(r)
        out = []
        for sig in (self.sigma_H1, self.sigma_He1, self.sigma_He2):
            out.append(self._integrate(b * sig))
        for sig, Eth in ((self.sigma_H1, ETH_H1), (self.sigma_He1, ETH_HE1),
                         (self.sigma_He2, ETH_HE2)):
            out.append(self._integrate(b * sig * (self.E - Eth) * EV_2_ERG))
        return tuple(out)

    def normalize_H1i(self, H1i, r=None):
        """Scale the shape so that the optically thin H I rate is ``H1i`` [1/s]."""
        self.shape = self.shape * (H1i / self.thin_rates(r)[0])


class BackgroundSource(Source):
    """Isotropic background; ``shape`` is Inu [erg/(s cm^2 Hz sr)]
    (``rad_src/background.py:249-325``)."""

    source_type = "background"

    def __init__(self, q_min, q_max, spectrum_type, Nnu=128, segmented=True,
                 px_fit_type="verner", verbose=False, z=None, T_eff=None, alpha=None,
                 user_E=None, user_shape=None):
        self._initialize(q_min, q_max, spectrum_type, Nnu, segmented, z, T_eff, alpha,
                         user_E, user_shape)

    @property
    def Inu(self):
        return self.shape

    def _base(self, r=None):
        if self.Nnu == 1:  # source_background.f90:111-112
            return 4.0 * PI * self.shape / self.E * ERG_2_EV
        return 4.0 * PI * self.shape / (self.E * H_ERG_S)

    def normalize_n(self, n):
        """photon number density [1/cm^3] (``background.py:597-630``)"""
        self.shape = self.shape * (n / (self._integrate(self._base()) / C_CM_S))


class PlaneSource(Source):
    """Plane-parallel source; ``shape`` is dFu/dnu [erg/(s cm^2 Hz)]
    (``rad_src/plane.py``).  An ``hm12`` plane source carries the flux of the
    isotropic background, 4 pi Inu (``plane.py:312-313``)."""

    source_type = "plane"

    def __init__(self, q_min, q_max, spectrum_type, Nnu=128, segmented=True,
                 px_fit_type="verner", verbose=False, z=None, T_eff=None, alpha=None,
                 user_E=None, user_shape=None):
        self._initialize(q_min, q_max, spectrum_type, Nnu, segmented, z, T_eff, alpha,
                         user_E, user_shape)
        if spectrum_type == "hm12":  # plane.py:312-313
            self.shape = self.shape * 4.0 * PI

    @property
    def dFu_over_dnu(self):
        return self.shape

    def _base(self, r=None):
        if self.Nnu == 1:  # source_plane.f90:101
            return self.shape / self.E * ERG_2_EV
        return self.shape / (self.E * H_ERG_S)

    def normalize_Fn(self, Fn):
        """photon flux [1/(s cm^2)]"""
        self.shape = self.shape * (Fn / self._integrate(self._base()))


class PointSource(Source):
    """Point source; ``shape`` is dLu/dnu [erg/(s Hz)] (``rad_src/point.py``)."""

    source_type = "point"

    def __init__(self, q_min, q_max, spectrum_type, Nnu=128, segmented=True,
                 px_fit_type="verner", verbose=False, z=None, T_eff=None, alpha=None,
                 user_E=None, user_shape=None):
        self._initialize(q_min, q_max, spectrum_type, Nnu, segmented, z, T_eff, alpha,
                         user_E, user_shape)

    @property
    def dLu_over_dnu(self):
        return self.shape

    def _base(self, r=None):
        dil = 1.0 if r is None else 1.0 / (4.0 * PI * r * r)
        if self.Nnu == 1:  # source_point.f90:104
            return self.shape / self.E * ERG_2_EV * dil
        return self.shape / (self.E * H_ERG_S) * dil

    def normalize_Ln(self, Ln):
        """photon luminosity [1/s] (``point.py:713-742``)"""
        self.shape = self.shape * (Ln / self._integrate(self._base()))
