"""rabacus_b200 -- B200-native (sm_100a, FP64 CUDA) hot path of galtay/rabacus.

Host code stays Python and mirrors the reference's public solver classes
(``SphereBgnd``, ``Slab2Bgnd``, ``SphereStromgren``, ``SlabPln``, ``Slab2Pln``);
the numerics run in ``librabacus_b200.so`` behind a C ABI
(``include/rabacus_b200.h``) that replaces the f2py module ``rabacus_fc``.
"""
from . import _lib  # noqa: F401
from ._lib import RabacusB200Error  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # lazy: the solver classes need numpy + the shared library
    import importlib
    if name in ("SphereBgnd", "SphereStromgren", "Slab2Bgnd", "SlabPln", "Slab2Pln",
                "Solve_CE", "Solve_PCE", "Solve_PCTE", "ChemistryRates", "CoolingRates"):
        return getattr(importlib.import_module(".solvers", __name__), name)
    if name in ("BackgroundSource", "PointSource", "PlaneSource"):
        return getattr(importlib.import_module(".rad_src", __name__), name)
    if name in ("rabacus_fc", "solvers", "rad_src", "plan", "configs", "build"):
        return importlib.import_module("." + name, __name__)
    if name == "Plan":
        return importlib.import_module(".plan", __name__).Plan
    raise AttributeError(name)
