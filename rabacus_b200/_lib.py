"""ctypes loader for ``librabacus_b200.so`` (the C ABI in ``include/rabacus_b200.h``).

This is where the reference loads its f2py extension ``rabacus_fc``
(``rabacus/f2py/sphere_bgnd.py:3``).  There is no CPU fallback: if the shared
library is missing, import of the compute layer fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RB_LIBRARY selects another build of the same sources (the checked build
# librabacus_b200_checked.so of build.py --checked: device assertions + buffer guards)
LIB_PATH = os.environ.get("RB_LIBRARY") or os.path.join(_HERE, "librabacus_b200.so")


class RabacusB200Error(RuntimeError):
    """Raised for every non-zero return code of the C ABI (the reference calls
    Fortran ``stop``: sphere_bgnd.f90:198-201, ion_solver.f90:536-546, ...)."""

    def __init__(self, code, msg):
        super().__init__("rabacus_b200 error %d: %s" % (code, msg))
        self.code = code


class rb_info(C.Structure):
    _fields_ = [("iters", C.c_int), ("n_launches", C.c_int), ("resid", C.c_double),
                ("ms_total", C.c_double), ("ms_device", C.c_double),
                ("ms_columns", C.c_double), ("ms_nu", C.c_double),
                ("ms_cell", C.c_double), ("ms_rec", C.c_double), ("ms_finish", C.c_double),
                ("evals", C.c_longlong)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class rb_plan_desc(C.Structure):
    _fields_ = [("geometry", C.c_int), ("n_problems", C.c_int), ("Nl", C.c_int),
                ("Nnu", C.c_int), ("Nmu", C.c_int), ("i_find_Teq", C.c_int),
                ("fixed_fcA", C.c_double), ("tol", C.c_double), ("device", C.c_int),
                ("max_sweeps", C.c_int), ("i_rec_meth", C.c_int), ("em_fac", C.c_double * 3),
                ("em_fac_set", C.c_int)]


GEOM_SPHERE_BGND, GEOM_SPHERE_STROMGREN, GEOM_SLAB_PLANE_1, GEOM_SLAB_PLANE_2, \
    GEOM_SLAB_BGND = range(5)

# every symbol declared in include/rabacus_b200.h
EXPORTS = (
    "rb_strerror", "rb_device_count", "rb_version", "rb_cache_clear",
    "rb_sphere_bgnd_solve", "rb_sphere_stromgren_solve", "rb_slab_bgnd_solve",
    "rb_slab_plane_solve", "rb_set_column_vs_impact",
    "rb_get_kchem", "rb_get_kcool", "rb_solve_ce", "rb_solve_pce", "rb_solve_pcte",
    "rb_photo_sigma", "rb_photo_Eth", "rb_gauss_legendre", "rb_photon_energies",
    "rb_plan_create", "rb_plan_destroy", "rb_plan_set_problem", "rb_plan_set_batch",
    "rb_plan_set_sphere_family", "rb_plan_set_sweep_order", "rb_plan_upload",
    "rb_plan_thin", "rb_plan_sweep_local", "rb_plan_finish_sweep", "rb_plan_finish_sweep_async",
    "rb_plan_poll_sweep", "rb_plan_solve",
    "rb_plan_get_problem", "rb_plan_get_rec", "rb_plan_get_all", "rb_plan_reserve_results", "rb_plan_get_all_pinned", "rb_plan_field_ptr", "rb_plan_stream", "rb_plan_sync",
    "rb_plan_evals_per_sweep", "rb_plan_last_info", "rb_plan_pack_cells",
    "rb_plan_unpack_cells",
    "rb_plan_peer_alloc", "rb_plan_peer_connect", "rb_plan_peer_deal", "rb_cell_of", "rb_plan_sweep_sharded", "rb_plan_solve_sharded",
)

# measurement / diagnostic symbols declared in rabacus_b200/csrc/rb_internal.h (not part of
# the drop-in boundary)
INTERNAL = ("rbi_diag_pow", "rbi_diag_hg97", "rbi_host_coefficients", "rbi_plan_set_libm_pow", "rbi_plan_set_far_field", "rbi_plan_bench_steps",
            "rbi_microbench", "rbi_last_zone_kernel_ms", "rbi_teq_table_fetch", "rbi_plan_fetch_inputs", "rbi_plan_check_guards", "rbi_diag_ratio_err")

_lib = None


def lib():
    """Load the shared library once; raise ImportError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). rabacus_b200 has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.rb_strerror.restype = C.c_char_p
        L.rb_strerror.argtypes = [C.c_int]
        L.rb_version.restype = C.c_char_p
        L.rb_photo_Eth.restype = C.c_double
        L.rb_photo_Eth.argtypes = [C.c_int]
        L.rb_plan_field_ptr.restype = C.c_void_p
        L.rb_plan_field_ptr.argtypes = [C.c_void_p, C.c_int]
        L.rb_plan_stream.restype = C.c_void_p
        L.rb_plan_stream.argtypes = [C.c_void_p]
        L.rb_plan_evals_per_sweep.restype = C.c_longlong
        L.rb_plan_evals_per_sweep.argtypes = [C.c_void_p]
        L.rb_plan_destroy.restype = None
        L.rbi_last_zone_kernel_ms.restype = C.c_double
        L.rb_cache_clear.restype = None
        L.rb_plan_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def check(code):
    if code != 0:
        raise RabacusB200Error(code, lib().rb_strerror(code).decode())


def device_count():
    return lib().rb_device_count()
