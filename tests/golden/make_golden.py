#!/usr/bin/env python
"""Generate the committed golden fixtures from the reference tree.

Run ONCE in the build container (``python tests/golden/make_golden.py``); the
GPU box has no /root/reference, so tests read only the files written here.

What is taken from the reference and how:

* ``hg97_py_twin.npz`` -- the 23 Hui & Gnedin 97 fits evaluated by the
  reference's *Python* coding ``rabacus/atomic/hui_gnedin_97.py`` (an
  implementation independent of the Fortran ``hui_gnedin_97.f90``; the
  reference's own test ``rabacus/unittest/test_atomic_rates.py:12`` requires the
  two to agree to 1e-10).  The module is Python-3 clean; it is imported with
  stub ``rabacus.constants`` modules because ``quantities`` is not installable.
* ``verner_py_twin.npz`` -- Verner+96 cross sections from the reference's Python
  coding ``rabacus/atomic/verner/photox/verner_photox.py:118-180`` (Python-2
  ``raise X, msg`` statements rewritten textually before exec) reading
  ``photo.dat``.
* ``hm12_spectrum_128.npz`` -- the 128-point HM12 z~3 spectrum fixture of the
  reference's profiling harness (``rabacus/f2py/profile/spectrum_E_eV.txt``,
  ``spectrum_Inu.txt``).
* ``hm12_photorates.npz`` -- the tabulated HM12 photo-rates
  (``rabacus/uv_bgnd/hm12_dat/photorates.out``), the pin used by
  ``rabacus/unittest/test_hm12.py``.
* ``../../rabacus_b200/data/hm12_uvb.npz`` -- the HM12 UVB table
  (``rabacus/uv_bgnd/hm12_dat/UVB.out``: published data of Haardt & Madau 2012),
  which the product's host-side spectrum builder needs.
"""
import os
import re
import sys
import types

import numpy as np

REF = os.environ.get("RABACUS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))


class Q(np.ndarray):
    """Minimal stand-in for quantities.Quantity: an ndarray with .units/.magnitude."""

    def __new__(cls, a):
        return np.asarray(a, dtype=np.float64).view(cls)

    @property
    def magnitude(self):
        return np.asarray(self)

    @property
    def units(self):
        return 1.0

    @units.setter
    def units(self, v):
        pass


def install_stubs():
    class Units:
        def __getattr__(self, name):
            return 1.0

    class PhysicalConstants:
        # rabacus/constants/physical.py:76-77 (kb in erg/K)
        kb = 1.3806488e-16

    rab = types.ModuleType("rabacus")
    const = types.ModuleType("rabacus.constants")
    phys = types.ModuleType("rabacus.constants.physical")
    units = types.ModuleType("rabacus.constants.units")
    utils_pkg = types.ModuleType("rabacus.utils")
    utils = types.ModuleType("rabacus.utils.utils")
    phys.PhysicalConstants = PhysicalConstants
    units.Units = Units

    class InputError(Exception):
        pass

    class NeedUnitsError(Exception):
        pass

    utils.InputError = InputError
    utils.NeedUnitsError = NeedUnitsError
    utils.isiterable = lambda x: hasattr(x, "__iter__") and np.ndim(x) > 0
    const.physical, const.units = phys, units
    utils_pkg.utils = utils
    rab.constants, rab.utils = const, utils_pkg
    sys.modules.update({
        "rabacus": rab, "rabacus.constants": const,
        "rabacus.constants.physical": phys, "rabacus.constants.units": units,
        "rabacus.utils": utils_pkg, "rabacus.utils.utils": utils,
    })


def load_py2_module(path, name):
    src = open(path).read()
    # `raise Err, 'msg'` -> `raise Err('msg')`
    src = re.sub(r"raise\s+([\w\.]+)\s*,\s*(.+)", r"raise \1(\2)", src)
    mod = types.ModuleType(name)
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


HG97_NAMES = ["reH2a", "reH2b", "reHe2a", "reHe2b", "reHe2di", "reHe3a", "reHe3b",
              "ciH1", "ciHe1", "ciHe2", "recH2a", "recH2b", "recHe2a", "recHe2b",
              "recHe2di", "recHe3a", "recHe3b", "cicH1", "cicHe1", "cicHe2",
              "cecH1", "cecHe2", "bremss"]


def main():
    install_stubs()

    # ---- HG97 python twin
    hg = load_py2_module(os.path.join(REF, "rabacus/atomic/hui_gnedin_97.py"),
                         "ref_hg97").HuiGnedin97()
    T = np.concatenate([np.linspace(1.0e4, 1.0e5, 128),      # test_atomic_rates.py grid
                        np.logspace(np.log10(7.5e3), 6.0, 64)])
    rates = np.stack([np.asarray(getattr(hg, n)(Q(T)), dtype=np.float64)
                      for n in HG97_NAMES], axis=1)
    np.savez(os.path.join(HERE, "hg97_py_twin.npz"), T=T, rates=rates,
             names=np.array(HG97_NAMES))
    print("hg97_py_twin.npz", rates.shape)

    # ---- Verner python twin
    vp = load_py2_module(
        os.path.join(REF, "rabacus/atomic/verner/photox/verner_photox.py"),
        "ref_verner")
    px = vp.PhotoXsections_Verner96()
    # make table columns unit-aware stand-ins
    for k in ("Eth", "Emax", "E0", "sigma0"):
        setattr(px, k, Q(np.asarray(getattr(px, k))))
    E = np.concatenate([13.6 * 10 ** np.linspace(0, np.log10(400.0), 257),
                        [13.6, 24.59, 54.42, 13.59999, 24.58999, 54.41999]])
    sig = np.stack([np.asarray(px.return_fit(Z, N, Q(E.copy())), dtype=np.float64)
                    for (Z, N) in ((1, 1), (2, 2), (2, 1))], axis=1)
    np.savez(os.path.join(HERE, "verner_py_twin.npz"), E=E, sigma=sig)
    print("verner_py_twin.npz", sig.shape)

    # ---- profile-harness spectrum
    E128 = np.loadtxt(os.path.join(REF, "rabacus/f2py/profile/spectrum_E_eV.txt"))
    I128 = np.loadtxt(os.path.join(REF, "rabacus/f2py/profile/spectrum_Inu.txt"))
    np.savez(os.path.join(HERE, "hm12_spectrum_128.npz"), E_eV=E128, Inu=I128)
    print("hm12_spectrum_128.npz", E128.shape)

    # ---- HM12 photorates table
    pr = np.loadtxt(os.path.join(REF, "rabacus/uv_bgnd/hm12_dat/photorates.out"))
    np.savez(os.path.join(HERE, "hm12_photorates.npz"), z=pr[:, 0], H1i=pr[:, 1],
             H1h_eV=pr[:, 2], He1i=pr[:, 3], He1h_eV=pr[:, 4], He2i=pr[:, 5],
             He2h_eV=pr[:, 6], comptonh_eV=pr[:, 7])
    print("hm12_photorates.npz", pr.shape)

    # ---- HM12 UVB table (product data)
    fname = os.path.join(REF, "rabacus/uv_bgnd/hm12_dat/UVB.out")
    with open(fname) as f:
        lines = f.readlines()
    z = np.array([float(v) for v in lines[20].split()])   # hm12.py:205-206
    dat = np.loadtxt(fname, skiprows=21)                   # hm12.py:208
    os.makedirs(os.path.join(REPO, "rabacus_b200", "data"), exist_ok=True)
    np.savez_compressed(os.path.join(REPO, "rabacus_b200", "data", "hm12_uvb.npz"),
                        z=z, lam_A=dat[:, 0], Inu=dat[:, 1:])
    print("hm12_uvb.npz", dat.shape, z.shape)


if __name__ == "__main__":
    main()
