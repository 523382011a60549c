"""Deterministic synthetic inputs for the configurations named in BASELINE.json
(SURVEY 8d).  All cgs, no RNG.  Pure host/numpy; shared by tests and bench.py.

Each builder returns a dict with the positional inputs of the corresponding
``rabacus_fc`` solve: Edges, nH, nHe, T, E_eV, shape plus flags.
"""
import numpy as np

from .rad_src import (ETH_H1, KPC_CM, BackgroundSource, PlaneSource, PointSource)

YP = 0.24


def hm12_background(Nnu, z=3.0):
    src = BackgroundSource(1.0, 400.0, "hm12", Nnu=Nnu, z=z)
    return src.E.copy(), src.shape.copy()


def c1_iliev_stromgren(Nl=128):
    """Iliev+06 test 1: monochromatic point source, nH=1e-3, T=1e4, 6.6 kpc
    (doc/guide_stromgren_sphere_examples.rst:119-125,160-167,207-210)."""
    src = PointSource(1.0, 1.0, "monochromatic")
    src.normalize_Ln(5.0e48)
    return dict(geometry="sphere_stromgren", Edges=np.linspace(0.0, 6.6 * KPC_CM, Nl + 1),
                nH=np.full(Nl, 1.0e-3), nHe=np.full(Nl, 1.0e-15), T=np.full(Nl, 1.0e4),
                E_eV=src.E.copy(), shape=src.shape.copy(), fixed_fcA=0.0, find_Teq=False,
                z=0.0, Nmu=0)


def c2_slab_plane(Nl=256, Nnu=128, find_Teq=False, two_sided=False):
    """SlabPln: L=150 kpc, nH=2.2e-3, HM12 plane source at z=3
    (doc/guide_slab_examples.rst:134-145,172-181)."""
    src = PlaneSource(1.0, 400.0, "hm12", Nnu=Nnu, z=3.0)
    nH = np.full(Nl, 2.2e-3)
    return dict(geometry="slab_plane_2" if two_sided else "slab_plane_1",
                Edges=np.linspace(0.0, 150.0 * KPC_CM, Nl + 1), nH=nH,
                nHe=nH * 0.25 * YP / (1.0 - YP), T=np.full(Nl, 1.0e4), E_eV=src.E.copy(),
                shape=src.shape.copy(), fixed_fcA=1.0, find_Teq=find_Teq, z=3.0, Nmu=0)


def c3_slab_bgnd(Nl=1024, Nnu=512, find_Teq=False):
    """Slab2Bgnd: the reference's profiling-harness slab
    (rabacus/f2py/profile/test_slab_bgnd.f90:46-62) under HM12 z=3."""
    E, shape = hm12_background(Nnu, 3.0)
    L = 1.3177850552122683e22
    return dict(geometry="slab_bgnd", Edges=np.arange(Nl + 1) * (L / Nl),
                nH=np.full(Nl, 8.218831761053482e-3), nHe=np.full(Nl, 6.765635287054012e-4),
                T=np.full(Nl, 1.0e4), E_eV=E, shape=shape, fixed_fcA=1.0, find_Teq=find_Teq,
                z=3.0, Nmu=0)


def c4_sphere_bgnd(Nl=4096, Nnu=512, Nmu=256, find_Teq=False, nH0=10.0, R_kpc=20.0, z=3.0,
                   spectrum=None):
    """SphereBgnd: R=20 kpc, nH = nH0 (r/r0)^-2 with a flat core r0=R/100, HM12 z=3
    (doc/guide_bgnd_sphere_examples.rst:126-146,205-231)."""
    E, shape = spectrum if spectrum is not None else hm12_background(Nnu, z)
    R = R_kpc * KPC_CM
    Edges = np.linspace(0.0, R, Nl + 1)
    r = 0.5 * (Edges[1:] + Edges[:-1])
    r0 = R / 100.0
    nH = nH0 * np.where(r < r0, 1.0, (r / r0) ** -2.0)
    return dict(geometry="sphere_bgnd", Edges=Edges, nH=nH, nHe=nH * 0.25 * YP / (1.0 - YP),
                T=np.full(Nl, 1.0e4), E_eV=E, shape=shape, fixed_fcA=1.0, find_Teq=find_Teq,
                z=z, Nmu=Nmu)


def c5_sweep(n_nH=32, n_R=16, n_z=16, Nl=512, Nnu=128, Nmu=32, find_Teq=False):
    """Parameter sweep of independent SphereBgnd problems:
    nH0 in logspace(-1,1.5), R in linspace(5,50) kpc, z in linspace(0,6)."""
    nH0s = np.logspace(-1.0, 1.5, n_nH)
    Rs = np.linspace(5.0, 50.0, n_R)
    zs = np.linspace(0.0, 6.0, n_z)
    spectra = [hm12_background(Nnu, float(z)) for z in zs]
    probs = []
    for nH0 in nH0s:
        for R in Rs:
            for iz, z in enumerate(zs):
                probs.append(c4_sphere_bgnd(Nl, Nnu, Nmu, find_Teq, nH0=float(nH0),
                                            R_kpc=float(R), z=float(z), spectrum=spectra[iz]))
    return probs


def c5_params(index, n_nH=32, n_R=16, n_z=16):
    """(nH0, R_kpc, z, iz) of problems ``index`` of :func:`c5_sweep` (same ordering:
    nH0 slowest, z fastest)."""
    index = np.asarray(index)
    nH0s = np.logspace(-1.0, 1.5, n_nH)
    Rs = np.linspace(5.0, 50.0, n_R)
    zs = np.linspace(0.0, 6.0, n_z)
    iz = index % n_z
    iR = (index // n_z) % n_R
    iH = index // (n_z * n_R)
    return nH0s[iH], Rs[iR], zs[iz], iz


def c5_batch_arrays(index, Nl=512, Nnu=128, n_nH=32, n_R=16, n_z=16):
    """The inputs of problems ``index`` of the C5 sweep as batch arrays, without per-problem
    Python work: the 16 spectra and the 16 density profiles (one per radius) are built once
    and the batch is assembled by broadcasting -- bit-identical to :func:`c4_sphere_bgnd`
    per problem.  Returns dict(Edges [n, Nl+1], nH, nHe, T [n, Nl], E, shape [n_z, Nnu],
    spec_index [n], z [n])."""
    nH0, R_kpc, z, iz = c5_params(index, n_nH, n_R, n_z)
    zs = np.linspace(0.0, 6.0, n_z)
    spectra = [hm12_background(Nnu, float(zz)) for zz in zs]
    Rs = np.linspace(5.0, 50.0, n_R)
    iR = (np.asarray(index) // n_z) % n_R
    edges_R = np.array([np.linspace(0.0, float(R) * KPC_CM, Nl + 1) for R in Rs])
    r = 0.5 * (edges_R[:, 1:] + edges_R[:, :-1])
    r0 = np.array([float(R) * KPC_CM / 100.0 for R in Rs])[:, None]
    unit = np.where(r < r0, 1.0, (r / r0) ** -2.0)          # per radius
    nH = nH0[:, None] * unit[iR]
    return dict(Edges=edges_R[iR], nH=nH, nHe=nH * 0.25 * YP / (1.0 - YP),
                T=np.full_like(nH, 1.0e4), E=np.array([s[0] for s in spectra]),
                shape=np.array([s[1] for s in spectra]), spec_index=iz.astype(np.int32), z=z)


def c5_stage(plan, index, n_nH=32, n_R=16, n_z=16, find_Teq=False):
    """Stage problems ``index`` of the C5 sweep into ``plan`` (slots 0 .. len(index)-1) with
    one ``rb_plan_set_batch`` call.  Returns a one-line description of the path taken."""
    a = c5_batch_arrays(index, plan.Nl, plan.Nnu, n_nH, n_R, n_z)
    plan.set_batch(0, a["Edges"], a["nH"], a["nHe"], a["T"], a["E"], a["shape"],
                   a["spec_index"], a["z"])
    return ("host: %d spectra + %d profiles built once (numpy), batch assembled by "
            "broadcasting, one rb_plan_set_batch call" % (n_z, n_R))


def c5_stage_device(plan, index, n_nH=32, n_R=16, n_z=16, find_Teq=False):
    """Stage problems ``index`` of the C5 sweep as a device-side family: only (nH0, R, z) per
    problem, the HM12 table and the energy grid go to the GPU; edges, density profiles,
    spectra, cross sections and per-frequency weights are produced there by
    ``plan.upload()`` (``rb_plan_set_sphere_family``).  Same problems as :func:`c5_stage`
    to the last few ulp (device ``pow`` / ``log10`` instead of numpy's)."""
    nH0, R_kpc, z, _ = c5_params(index, n_nH, n_R, n_z)
    plan.set_sphere_family(nH0, R_kpc * KPC_CM, z, core_div=100.0, slope=2.0,
                           nHe_over_nH=0.25 * YP / (1.0 - YP), T0=1.0e4, q_min=1.0, q_max=400.0)
    return ("device: (nH0, R, z) per problem + HM12 table uploaded, edges / profiles / spectra / "
            "cross sections / weights generated by k_family_profiles + k_family_spectra")


GEOM_IDS = {"sphere_bgnd": 0, "sphere_stromgren": 1, "slab_plane_1": 2, "slab_plane_2": 3,
            "slab_bgnd": 4}
