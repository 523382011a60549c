"""Post-solve reductions of the slab wrapper classes (host side, numpy):
mean molecular weight, central values and weighted averages
(``rabacus/f2py/slab_base.py:120-437``, ``rabacus/cosmology/jeans.py:45-64``)."""
import numpy as np

from .rad_src import trap

#: the fields ``calc_weighted`` averages (``slab_base.py:126-283``)
AVERAGED_FIELDS = (
    "xH1", "xH2", "xHe1", "xHe2", "xHe3", "T", "mu",
    "H1i", "H1i_src", "H1i_rec", "He1i", "He1i_src", "He1i_rec", "He2i", "He2i_src", "He2i_rec",
    "H1h", "H1h_src", "H1h_rec", "He1h", "He1h_src", "He1h_rec", "He2h", "He2h_src", "He2h_rec")
#: the fields sampled at the central layer (``slab_base.py:372-411``)
CENTRAL_FIELDS = ("T", "mu", "ne", "nH", "nHe") + AVERAGED_FIELDS[:5] + AVERAGED_FIELDS[7:]


class Averaged(object):
    """Attribute bag, as ``slab_base.Averaged``."""


class Central(object):
    """Attribute bag, as ``slab_base.Central``."""


def mean_molecular_weight(xH2, xHe2, xHe3, Yp=0.248):
    """Mean mass per particle in atomic mass units (``jeans.py:45-64``; the wrapper
    uses ``Jeans()`` with its default helium mass fraction 0.248, ``slab_base.py:302``)."""
    den = 4.0 - 3.0 * Yp + 4.0 * np.asarray(xH2) * (1.0 - Yp) \
        + (np.asarray(xHe2) + 2.0 * np.asarray(xHe3)) * Yp
    return 4.0 / den


def calc_weighted(obj, wa):
    """``<q>_w = trap(z_c, w q) / trap(z_c, w)`` for every averaged field of ``obj``
    (``slab_base.py:120-283``)."""
    wq = Averaged()
    den = trap(obj.z_c, wa)
    for name in AVERAGED_FIELDS:
        setattr(wq, name, trap(obj.z_c, wa * getattr(obj, name)) / den)
    return wq


def slab_averages(obj):
    """``obj.avg``: central-layer values and the six weightings (``slab_base.py:369-431``)."""
    avg = Averaged()
    cen = Central()
    Nl2 = obj.Nl // 2
    for name in CENTRAL_FIELDS:
        setattr(cen, name, getattr(obj, name)[Nl2])
    avg.cen = cen
    avg.v_w = calc_weighted(obj, obj.dz)
    avg.nH_w = calc_weighted(obj, obj.nH)
    avg.nHe_w = calc_weighted(obj, obj.nHe)
    avg.nH1_w = calc_weighted(obj, obj.nH * obj.xH1)
    avg.nHe1_w = calc_weighted(obj, obj.nHe * obj.xHe1)
    avg.nHe2_w = calc_weighted(obj, obj.nHe * obj.xHe2)
    return avg
