// rb_api.cu -- host side of librabacus_b200.so: plans, staging, iteration
// control and the extern "C" boundary declared in include/rabacus_b200.h.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#include <sys/mman.h>
#include <unistd.h>

#include "rb_rec.cuh"
#include "rb_inputs.cuh"
#include "rb_internal.h"

using namespace rb;

#define CK(call)                                   \
  do {                                             \
    cudaError_t e_ = (call);                       \
    if (e_ != cudaSuccess) {                       \
      rb_set_cuda_error(e_, __FILE__, __LINE__);   \
      return RB_ERR_CUDA;                          \
    }                                              \
  } while (0)

static thread_local char g_cuda_msg[256] = "";
static void rb_set_cuda_error(cudaError_t e, const char *file, int line) {
  snprintf(g_cuda_msg, sizeof(g_cuda_msg), "CUDA error: %s (%s:%d)", cudaGetErrorString(e),
           file, line);
}

extern "C" const char *rb_strerror(int code) {
  switch (code) {
    case RB_OK: return "ok";
    case RB_ERR_EDGES_NOT_MONOTONIC: return "Edges not monotonic";
    case RB_ERR_MAX_ITER_OUTER: return "MAX_ITER reached in the sweep loop";
    case RB_ERR_MAX_ITER_PCE: return "solve_pce: iter > MAX_ITER";
    case RB_ERR_MAX_ITER_PCTE: return "solve_pcte: iter > MAX_ITER";
    case RB_ERR_T_NOT_BRACKETED: return "Teq not bracketed";
    case RB_ERR_NAN: return "NaN in solve_ce";
    case RB_ERR_RAY_GEOMETRY: return "ray geometry undefined (no shell edge <= impact parameter)";
    case RB_ERR_BAD_ARG: return "bad argument";
    case RB_ERR_UNSUPPORTED: return "option not supported by the CUDA path (fits: verner / hg97 only)";
    case RB_ERR_CUDA: return g_cuda_msg[0] ? g_cuda_msg : "CUDA error / no CUDA device";
    case RB_ERR_NOMEM: return "out of memory";
    case RB_ERR_PEER_TIMEOUT: return "cell-sharded sweep: a peer rank did not reach the barrier within 2 s";
  }
  return "unknown error";
}

extern "C" const char *rb_version(void) { return "rabacus_b200 0.1 (sm_100a, fp64)"; }

extern "C" int rb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// ---------------------------------------------------------------------------
// host helpers: cross sections, Gauss-Legendre nodes, spectral tables
// ---------------------------------------------------------------------------
extern "C" int rb_photo_sigma(int species, const double *E_eV, int nn, double *sigma) {
  if (species < 0 || species > 2 || nn < 0) return RB_ERR_BAD_ARG;
  for (int i = 0; i < nn; ++i) sigma[i] = v96_sigma(species, E_eV[i]);
  return RB_OK;
}
extern "C" double rb_photo_Eth(int species) { return v96_row(species).Eth; }

// Gauss-Legendre nodes and weights from the Jacobi matrix of the Legendre
// recurrence (Golub & Welsch), diagonalised with the implicit QL algorithm --
// the same method as the reference (legendre_polynomial.f90:6-207,815-864), so
// nodes agree to rounding.  0-based arrays.
extern "C" int rb_gauss_legendre(int n, double *x, double *w) {
  if (n < 1) return RB_ERR_BAD_ARG;
  std::vector<double> d(n, 0.0), e(n + 1, 0.0), z(n, 0.0);
  for (int i = 1; i <= n; ++i) e[i - 1] = sqrt((double)(i * i) / (double)(4 * i * i - 1));
  z[0] = sqrt(2.0);
  const double eps = 2.220446049250313e-16;
  if (n > 1) {
    e[n - 1] = 0.0;
    for (int l = 0; l < n; ++l) {
      int iter = 0;
      for (;;) {
        int m = l;
        for (; m < n - 1; ++m)
          if (fabs(e[m]) <= eps * (fabs(d[m]) + fabs(d[m + 1]))) break;
        double p = d[l];
        if (m == l) break;
        if (iter++ >= 30) return RB_ERR_BAD_ARG;
        double g = (d[l + 1] - p) / (2.0 * e[l]);
        double r = sqrt(g * g + 1.0);
        g = d[m] - p + e[l] / (g + copysign(r, g));
        double s = 1.0, c = 1.0;
        p = 0.0;
        for (int i = m - 1; i >= l; --i) {
          double f = s * e[i];
          const double bb = c * e[i];
          if (fabs(g) <= fabs(f)) {
            c = g / f;
            r = sqrt(c * c + 1.0);
            e[i + 1] = f * r;
            s = 1.0 / r;
            c = c * s;
          } else {
            s = f / g;
            r = sqrt(s * s + 1.0);
            e[i + 1] = g * r;
            c = 1.0 / r;
            s = s * c;
          }
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * bb;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - bb;
          f = z[i + 1];
          z[i + 1] = s * z[i] + c * f;
          z[i] = c * z[i] - s * f;
        }
        d[l] = d[l] - p;
        e[l] = g;
        e[m] = 0.0;
      }
    }
    for (int i = 0; i < n - 1; ++i) {  // ascending selection sort, as the reference
      int k = i;
      double p = d[i];
      for (int j = i + 1; j < n; ++j)
        if (d[j] < p) { k = j; p = d[j]; }
      if (k != i) {
        d[k] = d[i]; d[i] = p;
        const double t = z[i]; z[i] = z[k]; z[k] = t;
      }
    }
  }
  for (int i = 0; i < n; ++i) { x[i] = d[i]; w[i] = z[i] * z[i]; }
  return RB_OK;
}

#ifdef RB_DEBUG_CHECKS
constexpr size_t kGuardBytes = 256;
#endif
struct DevBuf {
  double *d = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    n = count;
    if (cudaMalloc((void **)&d, std::max<size_t>(1, count) * 8) != cudaSuccess) {
      cudaGetLastError();
      d = nullptr;
      return RB_ERR_NOMEM;
    }
    return RB_OK;
  }
  ~DevBuf() { if (d) cudaFree(d); }
};

static int source_kind(int geom) {  // 0 background, 1 point, 2 plane
  switch (geom) {
    case RB_GEOM_SPHERE_BGND: case RB_GEOM_SLAB_BGND: return 0;
    case RB_GEOM_SPHERE_STROMGREN: return 1;
    default: return 2;
  }
}

// Per-frequency table for one problem: [Nnu][9] = sigma_X/sigma_X,th (3),
// W_ion,X (3), W_heat,X (3), plus thin[9] = sum_k W (6), min_nu sigma ratio (3).
//   y_ion,X(k)  = base_k * sigma_X(k)                  (source_background.f90:306-307)
//   y_heat,X(k) = y_ion,X(k) * (E_k - E_th,X) * eV_2_erg   (:350)
//   base_k = 4 pi I_k/(E_k h) | F_k/(E_k h) | L_k/(E_k h)   (s_bgnd:306, s_plane:296, s_point:348;
//            the point source's 1/(4 pi r^2) is applied per cell in the kernel)
//   trapezoid (utils.f90:19-20) folded into weights: W_k = y_k * (dE_{k-1}+dE_k)/2.
// Nnu == 1 (monochromatic): rate = Fn*sigma*a, heat = Fn*sigma*(E-Eth)*eV_2_erg*a
//   with Fn = 4 pi I/E*erg_2_eV | F/E*erg_2_eV | L/E*erg_2_eV
//   (source_background.f90:111-112,195-197,231-233).
static void build_spec_table(int geom, const double *E, const double *shape, int Nnu,
                             const double sigma_th[3], double *tab, double *thin) {
  const int kind = source_kind(geom);
  const double four_pi = 4.0 * RB_PI;
  for (int q = 0; q < kThinStride; ++q) thin[q] = 0.0;
  for (int k = 0; k < Nnu; ++k) {
    double *row = tab + (size_t)k * kSpecStride;
    double wk, base;
    if (Nnu == 1) {
      wk = 1.0;
      base = (kind == 0 ? four_pi * shape[0] : shape[0]) / E[0] * RB_ERG_2_EV;
    } else {
      const double dlo = (k > 0) ? (E[k] - E[k - 1]) : 0.0;
      const double dhi = (k < Nnu - 1) ? (E[k + 1] - E[k]) : 0.0;
      wk = 0.5 * (dlo + dhi);
      base = (kind == 0 ? four_pi * shape[k] : shape[k]) / (E[k] * RB_H_ERG_S);
    }
    for (int X = 0; X < 3; ++X) {
      const double sig = v96_sigma(X, E[k]);
      const double Eth = v96_row(X).Eth;
      row[X] = sig / sigma_th[X];                       // sphere_bgnd.f90:669-671
      const double yi = base * sig;
      const double yh = yi * (E[k] - Eth) * RB_EV_2_ERG;
      row[3 + X] = yi * wk;
      row[6 + X] = yh * wk;
      thin[X] += row[3 + X];
      thin[3 + X] += row[6 + X];
      // min_nu sigma_X(nu)/sigma_X,th: lower bound of a ray's optical depth (k_nu_rates)
      thin[6 + X] = (k == 0) ? row[X] : std::min(thin[6 + X], row[X]);
    }
  }
}

// ---------------------------------------------------------------------------
// Memo tables of the temperature bisection (TeqTab, rb_ion.cuh): one per (device, case-A
// fraction, libm flag, depth), shared by every plan and single-zone call of the process;
// built on first use (k_teq_table, 2^depth + 1 evaluations of the rate fits) and kept --
// the four most recently used ones -- until rb_cache_clear().
// Depth: RB_TEQ_DEPTH (0 disables the memo, default 18: 32 MiB, L2-resident hot levels).
// ---------------------------------------------------------------------------
struct TeqTable {
  int dev = 0, depth = 0, libm = 0;
  double fcA = 0.0;
  double *rows = nullptr;
  ~TeqTable() {
    if (rows) {
      int prev = -1;
      cudaGetDevice(&prev);
      cudaSetDevice(dev);
      cudaFree(rows);
      if (prev >= 0) cudaSetDevice(prev);
      cudaGetLastError();
    }
  }
};
static std::mutex g_teq_mu;
static std::vector<std::shared_ptr<TeqTable>> g_teq;   // most recently used last

static int teq_depth_wanted() {
  const char *e = getenv("RB_TEQ_DEPTH");
  int d = e ? atoi(e) : 18;
  return std::max(0, std::min(d, 24));
}

// the table of (current device, fcA, libm); nullptr when disabled or out of memory (the
// solvers then evaluate the fits at every probe, with identical results)
static std::shared_ptr<TeqTable> teq_table_get(double fcA, int libm) {
  const int depth = teq_depth_wanted();
  if (depth < 1) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  std::lock_guard<std::mutex> lk(g_teq_mu);
  for (size_t i = 0; i < g_teq.size(); ++i) {
    const auto &t = g_teq[i];
    if (t->dev == dev && t->depth == depth && t->libm == (libm != 0) &&
        memcmp(&t->fcA, &fcA, sizeof(double)) == 0) {
      auto hit = t;
      g_teq.erase(g_teq.begin() + i);
      g_teq.push_back(hit);
      return hit;
    }
  }
  auto t = std::make_shared<TeqTable>();
  t->dev = dev; t->depth = depth; t->libm = libm != 0; t->fcA = fcA;
  const size_t nrows = ((size_t)1 << depth) + 1;
  if (cudaMalloc((void **)&t->rows, nrows * kTeqStride * sizeof(double)) != cudaSuccess) {
    cudaGetLastError();
    t->rows = nullptr;
    return nullptr;
  }
  k_teq_table<<<(unsigned)((nrows + 127) / 128), 128>>>(t->rows, depth, fcA, t->libm);
  if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  g_teq.push_back(t);
  if (g_teq.size() > 4) g_teq.erase(g_teq.begin());
  return t;
}
static TeqTab teq_tab_of(const std::shared_ptr<TeqTable> &t) {
  TeqTab tb;
  tb.rows = t ? t->rows : nullptr;
  tb.depth = t ? t->depth : 0;
  tb.pad_ = 0;
  return tb;
}

// ---------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------
struct rb_plan {
  std::shared_ptr<TeqTable> teq;   // find_Teq: memo of the rate fits (bound by plan_bind_teq)
  rb_plan_desc d;
  int dev = 0;   // CUDA device the plan lives on (every entry point switches to it)
  int Ndir;
  double sigma_th[3];
  cudaStream_t stream = nullptr;
  // device buffers
  double *dEdges = nullptr, *dnH = nullptr, *dnHe = nullptr, *dT = nullptr;
  double *dx = nullptr;     // 5*B*Nl
  double *dsrc = nullptr;   // 6*B*Nl
  double *dconv = nullptr, *dkchem = nullptr;
  double4 *dpack = nullptr;
  double *dspec = nullptr, *dthin = nullptr, *dmu = nullptr, *dzHz = nullptr;
  double4 *dcol = nullptr;
  double *dfar = nullptr;       // kept far-field moment rows [B][Nl/S + 1][NT][3] (SphereBgnd)
  int far_slog = 3;             // S = 2^far_slog: every S-th row is kept (far_row_slog)
  bool use_far = false;         // columns: far field from the table, near field walked
  unsigned *ditems = nullptr;   // column-kernel work list (LPT order), see build_items
  int *dwork = nullptr;         // [B] next-item counters of the column kernel
  // recombination-photon transfer (rec_meth != fixed)
  int rec_meth = 1;
  double *drec = nullptr;       // 6*B*Nl output rates (always allocated; zero for fixed)
  double *drecbuf = nullptr;    // scratch: em[3], F0[3], dtau[3], r_c (10*B*Nl), tlo, records
  double4 *drecI = nullptr;     // isotropic on a plan whose col buffer is too small
  unsigned *diso_items = nullptr;  // isotropic: (32-cell block, ray) work list, longest first
  int n_iso_items = 0;
  std::vector<double> iso_edges;   // edges the list was built from
  int iso_cells[3] = {-1, -1, -1}; // (start, stride, n) of the cell list it was built for
  RecDev R;
  RecConst K;
  cudaEvent_t rev[kSlots];      // after the rec kernels of each in-flight sweep
  // CUDA graphs of one sweep (one per in-flight slot), see solve_loop
  cudaGraphExec_t gexec[kSlots] = {nullptr};
  int graph_nl[kSlots] = {0};          // kernel launches inside each graph
  // everything a captured sweep bakes in: the by-value kernel arguments (PlanDev, the
  // photon-transfer constants and scratch pointers), the method and the launch-shape knobs
  struct GraphKey {
    PlanDev P;
    RecConst K;
    RecDev R;
    PeerDev Q;
    int rec_meth, live, spec_g, cols_chunks, sharded, use_far, order, nu_ltab, far_split, cols_variant;
  };
  GraphKey gkey;
  bool gkey_valid = false;
  bool capturing = false, graphs_ok = true;
  int active_hint = 0;   // problems still iterating at the last read-back (0: all)
  int direct_live = -1;  // live_problems() of the last directly launched sweep
  bool slot_timed[kSlots] = {false};    // stage events of that slot are valid (direct launches)
  // cell sharding over peer memory
  PeerDev Q;                    // world == 1 when not connected
  void *dpeer = nullptr;        // this rank's flags (256 B) + staging
  void *peer_base[kMaxPeers] = {nullptr};  // mapped bases (own = dpeer)
  long long peer_sweeps = 0;    // sharded sweeps enqueued since rb_plan_peer_connect
  int peer_world = 1;
  int n_items = 0, items_start = -1, items_stride = -1, items_chunks = 1, n_sm = 148;
  bool items_far = false;
  double *dpart = nullptr;      // chunk partial sums of the column kernel (nchunk > 1)
  size_t part_bytes = 0;
  std::vector<double> items_edges;  // problem-0 edges the list was built from
  struct SpecEntry {                       // per-frequency tables of the last 32 distinct
    std::vector<double> E, shape, tab;     // spectra staged through rb_plan_set_problem
    double thin[kThinStride];
  };
  std::vector<SpecEntry> spec_cache;
  // device-side input producers (rb_plan_set_sphere_family): parameters of every problem,
  // the shared energy grid and the HM12 table live on the device; rb_plan_upload then
  // regenerates the inputs there instead of copying host staging
  std::vector<std::pair<const unsigned char *, const char *>> guards;   // checked build only
  int sweep_order = RB_ORDER_JACOBI;   // SphereBgnd: RB_ORDER_GAUSS_SEIDEL = the reference's
                                       // single-thread order (rb_plan_set_sweep_order)
  bool family = false;
  double *dfam = nullptr;       // nH0, R, z, Hz [4][B] | E [Nnu] | log_lam, z_tab, logI
  FamilyDev F;
  double *hall = nullptr;              // pinned [18][B][Nl] + iters: results of the whole batch
  double *dall = nullptr;              // its device-side gather buffer [18][B][Nl]
  size_t hall_bytes = 0;               // (rb_plan_reserve_results / rb_plan_get_all_pinned)
  unsigned char *dreversed = nullptr;  // [B] input orientation was flipped (rb_plan_get_all)
  int *dflags = nullptr;    // active, notconv, iters, err : 4*B, then nactive
  unsigned long long *dresid = nullptr;
  // pinned host staging
  double *hEdges = nullptr, *hnH = nullptr, *hnHe = nullptr, *hT = nullptr, *hspec = nullptr,
         *hthin = nullptr, *hzHz = nullptr, *hout = nullptr;
  int *hflags = nullptr;
  std::vector<char> reversed;
  std::vector<double> mu, wmu;
  cudaEvent_t ev[8];                 // [4],[5]: solve span; [6],[7]: bench step span
  cudaEvent_t sev[kSlots][5];        // per in-flight sweep: 4 stage boundaries + done
  bool ev_ok = false;
  long long sweeps_issued = 0, sweeps_polled = 0;
  int last_active = 0;
  rb_info info;
  PlanDev P;
};

static size_t cells(const rb_plan *p) { return (size_t)p->d.n_problems * p->d.Nl; }

// number of local cells of the deal (start, stride) over ncell swept cells (cell_of(),
// rb_common.cuh); -1: not a valid deal
static int deal_n_local(int ncell, int start, int stride) {
  if (start < 0 || stride == 0) return -1;
  if (stride > 0) return start < ncell ? (ncell - start + stride - 1) / stride : 0;
  const int W = -stride;
  if (start >= W || ncell % (32 * W) != 0) return -1;
  return ncell / W;
}

// find_Teq plans: point P.teq at the memo table of the plan's current (fcA, libm flag).
// Called at the start of every solve (rb_plan_thin): both can change on a cached plan.
// P is part of the graph key, so captured sweeps follow.
static void plan_bind_teq(rb_plan *p) {
  if (!p->d.i_find_Teq) { p->teq.reset(); p->P.teq = teq_tab_of(nullptr); return; }
  if (p->teq && p->teq->libm == (p->P.libm_pow != 0) && p->teq->depth == teq_depth_wanted() &&
      memcmp(&p->teq->fcA, &p->P.fcA, sizeof(double)) == 0) return;
  p->teq = teq_table_get(p->P.fcA, p->P.libm_pow);
  p->P.teq = teq_tab_of(p->teq);
}

// Every plan entry point runs on the plan's device and leaves the caller's current device
// as it found it (a process may hold plans on several GPUs, or share the thread with torch).
struct DevGuard {
  int prev = -1;
  bool switched = false;
  explicit DevGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    if (dev >= 0 && dev != prev) switched = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DevGuard() {
    if (switched && prev >= 0) cudaSetDevice(prev);
  }
};
#define PLAN_DEV(p) DevGuard dev_guard_((p) ? (p)->dev : -1)

static void destroy_graphs(rb_plan *p) {
  for (auto &g : p->gexec)
    if (g) { cudaGraphExecDestroy(g); g = nullptr; }
  p->gkey_valid = false;
}

extern "C" void rb_plan_destroy(rb_plan *p) {
  if (!p) return;
  PLAN_DEV(p);
  if (p->stream) cudaStreamSynchronize(p->stream);
  destroy_graphs(p);
  cudaFree(p->drec); cudaFree(p->drecbuf); cudaFree(p->drecI); cudaFree(p->diso_items);
  cudaFree(p->dpart); cudaFree(p->dreversed); cudaFree(p->dfar); cudaFree(p->dfam);
  if (p->hall) cudaFreeHost(p->hall);
  cudaFree(p->dall);
  if (p->dpeer) {
    for (int r = 0; r < kMaxPeers; ++r)
      if (p->peer_base[r] && p->peer_base[r] != p->dpeer) cudaIpcCloseMemHandle(p->peer_base[r]);
    cudaFree(p->dpeer);
  }
  cudaFree(p->dEdges); cudaFree(p->dnH); cudaFree(p->dnHe); cudaFree(p->dT);
  cudaFree(p->dx); cudaFree(p->dsrc); cudaFree(p->dconv); cudaFree(p->dkchem); cudaFree(p->dpack);
  cudaFree(p->dspec); cudaFree(p->dthin); cudaFree(p->dmu); cudaFree(p->dcol);
  cudaFree(p->dzHz); cudaFree(p->dflags); cudaFree(p->dresid); cudaFree(p->ditems); cudaFree(p->dwork);
  cudaFreeHost(p->hEdges); cudaFreeHost(p->hnH); cudaFreeHost(p->hnHe); cudaFreeHost(p->hT);
  cudaFreeHost(p->hspec); cudaFreeHost(p->hthin); cudaFreeHost(p->hzHz);
  cudaFreeHost(p->hout); cudaFreeHost(p->hflags);
  if (p->ev_ok) {
    for (auto &e : p->ev) cudaEventDestroy(e);
    for (auto &sl : p->sev) for (auto &e : sl) cudaEventDestroy(e);
    for (auto &e : p->rev) cudaEventDestroy(e);
  }
  if (p->stream) cudaStreamDestroy(p->stream);
  cudaGetLastError();
  delete p;
}

extern "C" int rb_plan_create(rb_plan **out, const rb_plan_desc *desc) {
  if (!out || !desc) return RB_ERR_BAD_ARG;
  *out = nullptr;
  const rb_plan_desc &d = *desc;
  if (d.geometry < 0 || d.geometry > RB_GEOM_SLAB_BGND || d.n_problems < 1 || d.Nl < 1 ||
      d.Nnu < 1 || !(d.tol > 0.0))
    return RB_ERR_BAD_ARG;
  if (d.geometry == RB_GEOM_SPHERE_BGND && d.Nmu < 1) return RB_ERR_BAD_ARG;
  const int rec_meth = d.i_rec_meth == 0 ? 1 : d.i_rec_meth;
  const bool sphere = d.geometry == RB_GEOM_SPHERE_BGND || d.geometry == RB_GEOM_SPHERE_STROMGREN;
  if (rec_meth < 1 || rec_meth > (sphere ? 4 : 2)) return RB_ERR_BAD_ARG;
  if (rec_meth == 4 && d.Nmu < 1) return RB_ERR_BAD_ARG;
  if (d.Nl > 512 * 16) return RB_ERR_UNSUPPORTED;  // lock-step thin-state kernel limit
  if (rb_device_count() < 1) {
    snprintf(g_cuda_msg, sizeof(g_cuda_msg), "no CUDA device visible (there is no CPU fallback)");
    return RB_ERR_CUDA;
  }
  int dev_now = 0;
  if (cudaGetDevice(&dev_now) != cudaSuccess) cudaGetLastError();
  const int dev = d.device >= 0 ? d.device : dev_now;
  DevGuard dev_guard_(dev);
  if (dev != dev_now && !dev_guard_.switched) {
    snprintf(g_cuda_msg, sizeof(g_cuda_msg), "cannot select CUDA device %d", dev);
    return RB_ERR_CUDA;
  }
  rb_plan *p = new rb_plan();
  p->d = d;
  p->d.device = dev;
  p->dev = dev;
  switch (d.geometry) {
    case RB_GEOM_SPHERE_BGND: p->Ndir = d.Nmu; break;
    case RB_GEOM_SLAB_PLANE_2: case RB_GEOM_SLAB_BGND: p->Ndir = 2; break;
    default: p->Ndir = 1;
  }
  for (int X = 0; X < 3; ++X) p->sigma_th[X] = v96_sigma(X, v96_row(X).Eth);
  const size_t B = d.n_problems, Nl = d.Nl, Nnu = d.Nnu, NC = B * Nl;
#ifdef RB_DEBUG_CHECKS
  // checked build: kGuardBytes of 0xA5 behind every device buffer (rbi_plan_check_guards)
#define ALLOC(ptr, n)                                                          \
  do {                                                                         \
    const size_t n_ = ((size_t)(n) + 15) / 16 * 16;                            \
    if (cudaMalloc((void **)&(ptr), n_ + kGuardBytes) != cudaSuccess) {        \
      cudaGetLastError();                                                      \
      rb_plan_destroy(p);                                                      \
      return RB_ERR_NOMEM;                                                     \
    }                                                                          \
    cudaMemset((char *)(ptr) + n_, 0xA5, kGuardBytes);                         \
    p->guards.emplace_back((const unsigned char *)(ptr) + n_, #ptr);           \
  } while (0)
#else
#define ALLOC(ptr, n)                                                          \
  do {                                                                         \
    if (cudaMalloc((void **)&(ptr), (n)) != cudaSuccess) {                     \
      cudaGetLastError();                                                      \
      rb_plan_destroy(p);                                                      \
      return RB_ERR_NOMEM;                                                     \
    }                                                                          \
  } while (0)
#endif
#define HALLOC(ptr, n)                                                         \
  do {                                                                         \
    if (cudaMallocHost((void **)&(ptr), (n)) != cudaSuccess) {                 \
      cudaGetLastError();                                                      \
      rb_plan_destroy(p);                                                      \
      return RB_ERR_NOMEM;                                                     \
    }                                                                          \
  } while (0)
  ALLOC(p->dEdges, B * (Nl + 1) * 8); ALLOC(p->dnH, NC * 8); ALLOC(p->dnHe, NC * 8);
  ALLOC(p->dT, NC * 8); ALLOC(p->dx, 5 * NC * 8); ALLOC(p->dsrc, 6 * NC * 8);
  ALLOC(p->dconv, NC * 8); ALLOC(p->dkchem, 6 * NC * 8); ALLOC(p->dpack, B * (Nl + 1 + kPackPad) * sizeof(double4));
  ALLOC(p->dspec, B * Nnu * kSpecStride * 8); ALLOC(p->dthin, B * kThinStride * 8);
  ALLOC(p->dmu, 2 * (size_t)std::max(1, d.Nmu) * 8);
  ALLOC(p->dcol, NC * (size_t)p->Ndir * sizeof(double4)); ALLOC(p->dzHz, B * 2 * 8);
  ALLOC(p->dwork, B * sizeof(int));
  ALLOC(p->drec, 6 * NC * 8);
  p->rec_meth = rec_meth;
  if (rec_meth != 1) {
    // em[3] F0[3] dtau[3] r_c | tlo [B][3][Nl+1] | recA, recB [B][Nl+1] double4
    const size_t n_scal = 10 * NC + 3 * B * (Nl + 1);
    ALLOC(p->drecbuf, n_scal * 8 + 2 * B * (Nl + 1) * sizeof(double4) + 32);
    if (rec_meth == 4 && d.geometry != RB_GEOM_SPHERE_BGND)
      ALLOC(p->drecI, NC * (size_t)d.Nmu * sizeof(double4));
    if (rec_meth == 4) {
      ALLOC(p->diso_items, (size_t)((Nl + 31) / 32) * d.Nmu * sizeof(unsigned));
    }
  }
  if (d.geometry == RB_GEOM_SPHERE_BGND)
    ALLOC(p->ditems, (size_t)((Nl + 31) / 32) * ((d.Nmu + 1) / 2) * 8 * sizeof(unsigned));
  if (d.geometry == RB_GEOM_SPHERE_BGND) {
    // far-field moment rows: where shell records + kept rows fit in shared memory (the
    // column kernel that uses them); RB_NO_FAR=1 disables the far field
    p->far_slog = far_row_slog((int)Nl);
    const size_t far_bytes = B * (size_t)far_rows((int)Nl, p->far_slog) * kFarRow * sizeof(double);
    const int nseg = (int)((Nl + 1 + kFarSeg - 1) / kFarSeg);
    const char *nf = getenv("RB_NO_FAR");
    if (!(nf && atoi(nf) != 0) && far_cols_smem((int)Nl, p->far_slog) <= (size_t)226 * 1024 &&
        nseg <= 8 * 42) {
      if (cudaMalloc((void **)&p->dfar, far_bytes) == cudaSuccess) p->use_far = true;
      else { cudaGetLastError(); p->dfar = nullptr; }
    }
  }
  cudaDeviceGetAttribute(&p->n_sm, cudaDevAttrMultiProcessorCount, dev);
  ALLOC(p->dflags, (3 * B + kSlots) * sizeof(int)); ALLOC(p->dresid, B * 8);
  ALLOC(p->dreversed, B);
  HALLOC(p->hEdges, B * (Nl + 1) * 8); HALLOC(p->hnH, NC * 8); HALLOC(p->hnHe, NC * 8);
  HALLOC(p->hT, NC * 8); HALLOC(p->hspec, B * Nnu * kSpecStride * 8);
  HALLOC(p->hthin, B * kThinStride * 8); HALLOC(p->hzHz, B * 2 * 8); HALLOC(p->hout, 12 * Nl * 8);
  HALLOC(p->hflags, (3 * B + kSlots) * sizeof(int));
#undef ALLOC
#undef HALLOC
  p->reversed.assign(B, 0);
  if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess) {
    rb_plan_destroy(p);
    return RB_ERR_CUDA;
  }
  for (auto &e : p->ev) cudaEventCreate(&e);
  for (auto &sl : p->sev) for (auto &e : sl) cudaEventCreate(&e);
  for (auto &e : p->rev) cudaEventCreate(&e);
  p->ev_ok = true;
  if (d.geometry == RB_GEOM_SPHERE_BGND || rec_meth == 4) {
    p->mu.resize(d.Nmu); p->wmu.resize(d.Nmu);
    rb_gauss_legendre(d.Nmu, p->mu.data(), p->wmu.data());  // sphere_bgnd.f90:743
    cudaMemcpy(p->dmu, p->mu.data(), d.Nmu * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(p->dmu + d.Nmu, p->wmu.data(), d.Nmu * 8, cudaMemcpyHostToDevice);
  }
  memset(&p->info, 0, sizeof(p->info));
  memset(&p->Q, 0, sizeof(p->Q));
  p->Q.world = 1;
  PlanDev &P = p->P;
  memset(&P, 0, sizeof(P));
  P.B = d.n_problems; P.Nl = d.Nl; P.Nnu = d.Nnu; P.Nmu = d.Nmu; P.Ndir = p->Ndir;
  P.geom = d.geometry; P.find_Teq = d.i_find_Teq; P.max_sweeps = d.max_sweeps;
  P.libm_pow = 0; P.pad0_ = 0;
  P.fcA = d.fixed_fcA; P.tol = d.tol; P.tol_inner = d.tol * RB_TOL_FRAC;
  for (int X = 0; X < 3; ++X) P.sigma_th[X] = p->sigma_th[X];
  P.Edges = p->dEdges; P.nH = p->dnH; P.nHe = p->dnHe; P.T = p->dT;
  for (int i = 0; i < 5; ++i) P.x[i] = p->dx + i * NC;
  for (int i = 0; i < 6; ++i) P.src[i] = p->dsrc + i * NC;
  P.conv_old = p->dconv; P.kchem = p->dkchem; P.pack = p->dpack; P.spec = p->dspec; P.thin = p->dthin;
  P.mu = p->dmu; P.wmu = p->dmu + std::max(1, d.Nmu); P.col = p->dcol; P.zHz = p->dzHz;
  if (d.geometry == RB_GEOM_SPHERE_BGND || rec_meth == 4) P.wmu = p->dmu + d.Nmu;
  for (int i = 0; i < 6; ++i) P.rec[i] = p->drec + i * NC;
  if (rec_meth != 1) P.fcA = 1.0;  // sphere_bgnd.f90:435-443,676-684: case A with photon transfer
  {
    RecDev &R = p->R;
    memset(&R, 0, sizeof(R));
    for (int i = 0; i < 6; ++i) R.rec[i] = P.rec[i];
    if (rec_meth != 1) {
      double *q = p->drecbuf;
      for (int X = 0; X < 3; ++X) { R.em[X] = q + X * NC; R.F0[X] = q + (3 + X) * NC;
                                    R.dtau[X] = q + (6 + X) * NC; }
      R.r_c = q + 9 * NC;
      R.tlo = q + 10 * NC;
      size_t off = (10 * NC + 3 * B * (Nl + 1)) * 8;
      off = (off + 31) / 32 * 32;
      R.recA = (double4 *)((char *)p->drecbuf + off);
      R.recB = R.recA + B * (Nl + 1);
      R.Iray = p->drecI ? p->drecI : p->dcol;  // SphereBgnd: the ray buffer is free until
                                               // the column kernel of the same sweep
    }
    RecConst &K = p->K;
    // sphere_base.f90:224-236
    K.sH1_H1 = p->sigma_th[0]; K.sHe1_He1 = p->sigma_th[1]; K.sHe2_He2 = p->sigma_th[2];
    for (int X = 0; X < 3; ++X) K.E_th[X] = v96_row(X).Eth;
    K.sH1_He1 = v96_sigma(0, K.E_th[1]);
    K.sH1_He2 = v96_sigma(0, K.E_th[2]);
    K.sHe1_He2 = v96_sigma(1, K.E_th[2]);
    // sphere_base.f90:955: the factors multiply the emissivities, so explicit zeros mean
    // "no recombination emission"; only a desc that never set them (em_fac_set == 0) gets 1
    for (int X = 0; X < 3; ++X) K.em_fac[X] = d.em_fac_set ? d.em_fac[X] : 1.0;
  }
  P.active = p->dflags; P.iters = p->dflags + B; P.err = p->dflags + 2 * B;
  P.nactive = p->dflags + 3 * B; P.resid_bits = p->dresid;
  *out = p;
  return RB_OK;
}

static bool mono_inc(const double *a, int n) {  // sphere_base.f90:1493-1509
  for (int i = 1; i < n; ++i) if (a[i] < a[i - 1]) return false;
  return true;
}
static bool mono_dec(const double *a, int n) {  // sphere_base.f90:1467-1483
  for (int i = 1; i < n; ++i) if (a[i] > a[i - 1]) return false;
  return true;
}

extern "C" int rb_plan_set_problem(rb_plan *p, int b, const double *Edges, const double *nH,
                                   const double *nHe, const double *T, const double *E_eV,
                                   const double *shape, double z, double Hz) {
  if (!p || b < 0 || b >= p->d.n_problems) return RB_ERR_BAD_ARG;
  const int Nl = p->d.Nl, Nnu = p->d.Nnu, g = p->d.geometry;
  p->family = false;   // host staging takes over again
  // orientation: SphereBgnd wants decreasing Edges (sphere_bgnd.f90:182-202),
  // SphereStromgren increasing (sphere_stromgren.f90:186-204); slabs as given.
  bool rev = false;
  if (g == RB_GEOM_SPHERE_BGND) {
    if (mono_inc(Edges, Nl + 1)) rev = true;
    else if (!mono_dec(Edges, Nl + 1)) return RB_ERR_EDGES_NOT_MONOTONIC;
  } else if (g == RB_GEOM_SPHERE_STROMGREN) {
    if (mono_dec(Edges, Nl + 1)) rev = true;
    else if (!mono_inc(Edges, Nl + 1)) return RB_ERR_EDGES_NOT_MONOTONIC;
  }
  p->reversed[b] = rev;
  double *hE = p->hEdges + (size_t)b * (Nl + 1);
  double *hH = p->hnH + (size_t)b * Nl, *hHe = p->hnHe + (size_t)b * Nl,
         *hT = p->hT + (size_t)b * Nl;
  for (int i = 0; i <= Nl; ++i) hE[i] = Edges[rev ? Nl - i : i];
  for (int i = 0; i < Nl; ++i) {
    const int s = rev ? Nl - 1 - i : i;
    hH[i] = nH[s]; hHe[i] = nHe[s]; hT[i] = T[s];
  }
  // parameter sweeps reuse a handful of spectra over thousands of problems (C5: 16 for
  // 8192): the per-frequency table (3 Verner cross sections per frequency) is built once
  // per distinct spectrum (the last 32 are kept)
  double *tab = p->hspec + (size_t)b * Nnu * kSpecStride, *thin = p->hthin + (size_t)b * kThinStride;
  int hit = -1;
  for (size_t q = 0; q < p->spec_cache.size() && hit < 0; ++q) {
    const rb_plan::SpecEntry &e = p->spec_cache[q];
    if (memcmp(e.E.data(), E_eV, sizeof(double) * Nnu) == 0 &&
        memcmp(e.shape.data(), shape, sizeof(double) * Nnu) == 0)
      hit = (int)q;
  }
  if (hit >= 0) {
    memcpy(tab, p->spec_cache[hit].tab.data(), sizeof(double) * Nnu * kSpecStride);
    memcpy(thin, p->spec_cache[hit].thin, sizeof(double) * kThinStride);
  } else {
    build_spec_table(g, E_eV, shape, Nnu, p->sigma_th, tab, thin);
    if (p->spec_cache.size() >= 32) p->spec_cache.erase(p->spec_cache.begin());
    p->spec_cache.emplace_back();
    rb_plan::SpecEntry &e = p->spec_cache.back();
    e.E.assign(E_eV, E_eV + Nnu);
    e.shape.assign(shape, shape + Nnu);
    e.tab.assign(tab, tab + (size_t)Nnu * kSpecStride);
    memcpy(e.thin, thin, sizeof(double) * kThinStride);
  }
  p->hzHz[2 * b] = z;
  p->hzHz[2 * b + 1] = Hz;
  return RB_OK;
}

// Stage n problems b0 .. b0+n-1 in one call (parameter sweeps: thousands of problems that
// share a handful of spectra).  Row-major inputs; spec_index[i] selects problem i's
// spectrum among the nspec rows of E_eV / shape.
extern "C" int rb_plan_set_batch(rb_plan *p, int b0, int n, const double *Edges,
                                 const double *nH, const double *nHe, const double *T,
                                 int nspec, const double *E_eV, const double *shape,
                                 const int *spec_index, const double *z, const double *Hz) {
  if (!p || b0 < 0 || n < 0 || b0 + n > p->d.n_problems || nspec < 1 || !Edges || !nH || !nHe ||
      !T || !E_eV || !shape || !spec_index || !z)
    return RB_ERR_BAD_ARG;
  const size_t Nl = p->d.Nl, Nnu = p->d.Nnu;
  for (int i = 0; i < n; ++i) {
    const int q = spec_index[i];
    if (q < 0 || q >= nspec) return RB_ERR_BAD_ARG;
    const int rc = rb_plan_set_problem(p, b0 + i, Edges + (size_t)i * (Nl + 1), nH + i * Nl,
                                       nHe + i * Nl, T + i * Nl, E_eV + (size_t)q * Nnu,
                                       shape + (size_t)q * Nnu, z[i], Hz ? Hz[i] : 0.0);
    if (rc) return rc;
  }
  return RB_OK;
}

// Photon-energy grid of rad_src/source.py:178-259 (segmented): Nnu log-spaced samples of
// q = E / 13.6 eV between q_min and q_max, the samples nearest the three ionisation
// thresholds snapped onto them.  numpy.linspace semantics (start + k*step, exact end point).
static int photon_energies(double q_min, double q_max, int Nnu, std::vector<double> &E) {
  if (Nnu < 2 || !(q_min > 0.0) || !(q_max > q_min)) return RB_ERR_BAD_ARG;
  const double Eth[3] = {v96_row(0).Eth, v96_row(1).Eth, v96_row(2).Eth};
  const double a = log10(q_min), b = log10(q_max);
  const double step = (b - a) / (double)(Nnu - 1);
  std::vector<double> lq(Nnu), q(Nnu);
  for (int k = 0; k < Nnu; ++k) lq[k] = (k == Nnu - 1) ? b : (double)k * step + a;
  for (int k = 0; k < Nnu; ++k) q[k] = pow(10.0, lq[k]);
  double qmin = q[0], qmax = q[0];
  for (int k = 1; k < Nnu; ++k) { qmin = std::min(qmin, q[k]); qmax = std::max(qmax, q[k]); }
  const double qt[3] = {1.0, Eth[1] / Eth[0], Eth[2] / Eth[0]};
  if (qmin > 1.0 || qmax < qt[2]) return RB_ERR_BAD_ARG;   // source.py:205-212
  for (int X = 0; X < 3; ++X) {
    const double t = log10(qt[X]);
    int best = 0;
    for (int k = 1; k < Nnu; ++k)
      if (fabs(lq[k] - t) < fabs(lq[best] - t)) best = k;
    q[best] = qt[X];
  }
  E.resize(Nnu);
  for (int k = 0; k < Nnu; ++k) E[k] = q[k] * Eth[0];
  return RB_OK;
}

extern "C" int rb_photon_energies(double q_min, double q_max, int Nnu, double *E_eV) {
  if (!E_eV) return RB_ERR_BAD_ARG;
  std::vector<double> E;
  const int rc = photon_energies(q_min, q_max, Nnu, E);
  if (rc) return rc;
  memcpy(E_eV, E.data(), (size_t)Nnu * 8);
  return RB_OK;
}

extern "C" int rb_plan_set_sphere_family(rb_plan *p, const double *nH0, const double *R_cm,
                                         const double *z, const double *Hz, double core_div,
                                         double slope, double nHe_over_nH, double T0,
                                         double q_min, double q_max, int nlam, int nzt,
                                         const double *log_lam_cm, const double *z_tab,
                                         const double *logInu, double *E_eV_out) {
  if (!p || !nH0 || !R_cm || !z || !log_lam_cm || !z_tab || !logInu || nlam < 2 || nzt < 2 ||
      !(core_div > 0.0))
    return RB_ERR_BAD_ARG;
  if (p->d.geometry != RB_GEOM_SPHERE_BGND) return RB_ERR_UNSUPPORTED;
  PLAN_DEV(p);
  const size_t B = p->d.n_problems, Nl = p->d.Nl, Nnu = p->d.Nnu;
  for (size_t b = 0; b < B; ++b)
    if (!(R_cm[b] > 0.0) || !(nH0[b] >= 0.0)) return RB_ERR_BAD_ARG;
  for (int i = 1; i < nlam; ++i)
    if (!(log_lam_cm[i] >= log_lam_cm[i - 1])) return RB_ERR_BAD_ARG;   // repeated nodes allowed
  for (int i = 1; i < nzt; ++i)
    if (!(z_tab[i] > z_tab[i - 1])) return RB_ERR_BAD_ARG;
  std::vector<double> E;
  int rc = photon_energies(q_min, q_max, (int)Nnu, E);
  if (rc) return rc;
  if (E_eV_out) memcpy(E_eV_out, E.data(), Nnu * 8);
  const size_t n = 4 * B + Nnu + (size_t)nlam + (size_t)nzt + (size_t)nlam * nzt;
  std::vector<double> h(n);
  double *q = h.data();
  memcpy(q, nH0, B * 8); memcpy(q + B, R_cm, B * 8); memcpy(q + 2 * B, z, B * 8);
  for (size_t b = 0; b < B; ++b) q[3 * B + b] = Hz ? Hz[b] : 0.0;
  q += 4 * B;
  memcpy(q, E.data(), Nnu * 8); q += Nnu;
  memcpy(q, log_lam_cm, (size_t)nlam * 8); q += nlam;
  memcpy(q, z_tab, (size_t)nzt * 8); q += nzt;
  memcpy(q, logInu, (size_t)nlam * nzt * 8);
  cudaFree(p->dfam);
  p->dfam = nullptr;
  if (cudaMalloc((void **)&p->dfam, n * 8) != cudaSuccess) { cudaGetLastError(); return RB_ERR_NOMEM; }
  CK(cudaMemcpyAsync(p->dfam, h.data(), n * 8, cudaMemcpyHostToDevice, p->stream));
  CK(cudaStreamSynchronize(p->stream));   // `h` is pageable
  FamilyDev &F = p->F;
  memset(&F, 0, sizeof(F));
  F.B = (int)B; F.Nl = (int)Nl; F.Nnu = (int)Nnu;
  F.nH0 = p->dfam; F.R = p->dfam + B; F.z = p->dfam + 2 * B; F.Hz = p->dfam + 3 * B;
  F.E = p->dfam + 4 * B;
  F.core_div = core_div; F.slope = slope; F.he_ratio = nHe_over_nH; F.T0 = T0;
  F.hm.nlam = nlam; F.hm.nz = nzt;
  F.hm.log_lam = F.E + Nnu; F.hm.z = F.hm.log_lam + nlam; F.hm.logI = F.hm.z + nzt;
  F.Edges = p->dEdges; F.nH = p->dnH; F.nHe = p->dnHe; F.T = p->dT; F.spec = p->dspec;
  F.thin = p->dthin; F.zHz = p->dzHz; F.reversed = p->dreversed;
  // host mirrors that the launch logic reads: problem 0's edges (work-list order of the
  // column kernel), the orientation flags (results go back centre-first)
  for (size_t i = 0; i <= Nl; ++i) {
    const size_t k = Nl - i;
    p->hEdges[i] = (k == Nl) ? R_cm[0] : (double)k * (R_cm[0] / (double)Nl);
  }
  std::fill(p->reversed.begin(), p->reversed.end(), (char)1);
  p->family = true;
  return RB_OK;
}

extern "C" int rbi_plan_fetch_inputs(rb_plan *p, int b, double *Edges, double *nH, double *nHe,
                                     double *T, double *spec, double *thin) {
  if (!p || b < 0 || b >= p->d.n_problems) return RB_ERR_BAD_ARG;
  PLAN_DEV(p);
  const size_t Nl = p->d.Nl, Nnu = p->d.Nnu;
  CK(cudaStreamSynchronize(p->stream));
  if (Edges) CK(cudaMemcpy(Edges, p->dEdges + b * (Nl + 1), (Nl + 1) * 8, cudaMemcpyDeviceToHost));
  if (nH) CK(cudaMemcpy(nH, p->dnH + b * Nl, Nl * 8, cudaMemcpyDeviceToHost));
  if (nHe) CK(cudaMemcpy(nHe, p->dnHe + b * Nl, Nl * 8, cudaMemcpyDeviceToHost));
  if (T) CK(cudaMemcpy(T, p->dT + b * Nl, Nl * 8, cudaMemcpyDeviceToHost));
  if (spec) CK(cudaMemcpy(spec, p->dspec + b * Nnu * kSpecStride, Nnu * kSpecStride * 8,
                          cudaMemcpyDeviceToHost));
  if (thin) CK(cudaMemcpy(thin, p->dthin + b * kThinStride, kThinStride * 8, cudaMemcpyDeviceToHost));
  return RB_OK;
}

extern "C" int rb_plan_upload(rb_plan *p) {
  if (!p) return RB_ERR_BAD_ARG;
  PLAN_DEV(p);
  const size_t B = p->d.n_problems, Nl = p->d.Nl, Nnu = p->d.Nnu, NC = B * Nl;
  cudaStream_t s = p->stream;
  plan_bind_teq(p);
  if (p->family) {
    // inputs produced on the device from (nH0, R, z) per problem (rb_inputs.cuh)
    dim3 gp((unsigned)((Nl + 1 + 127) / 128), (unsigned)B);
    k_family_profiles<<<gp, 128, 0, s>>>(p->F);
    k_family_spectra<<<(unsigned)B, 128, 0, s>>>(p->F, p->sigma_th[0], p->sigma_th[1],
                                                 p->sigma_th[2]);
    p->info.n_launches += 2;
    CK(cudaGetLastError());
  } else {
    CK(cudaMemcpyAsync(p->dEdges, p->hEdges, B * (Nl + 1) * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->dnH, p->hnH, NC * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->dnHe, p->hnHe, NC * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->dT, p->hT, NC * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->dspec, p->hspec, B * Nnu * kSpecStride * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->dthin, p->hthin, B * kThinStride * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->dzHz, p->hzHz, B * 2 * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->dreversed, p->reversed.data(), B, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));  // `reversed` is pageable
  }
  for (size_t b = 0; b < B; ++b) {
    p->hflags[b] = 1; p->hflags[B + b] = 0; p->hflags[2 * B + b] = 0;
  }
  for (int q = 0; q < kSlots; ++q) p->hflags[3 * B + q] = 0;
  CK(cudaMemcpyAsync(p->dflags, p->hflags, (3 * B + kSlots) * sizeof(int), cudaMemcpyHostToDevice, s));
  p->sweeps_issued = p->sweeps_polled = 0;
  p->active_hint = 0;
  p->direct_live = -1;
  p->last_active = (int)B;
  CK(cudaMemsetAsync(p->dresid, 0, B * 8, s));
  memset(&p->info, 0, sizeof(p->info));
  p->info.resid = NAN;
  return RB_OK;
}

extern "C" int rb_plan_thin(rb_plan *p) {
  if (!p) return RB_ERR_BAD_ARG;
  PLAN_DEV(p);
  const int Nl = p->d.Nl, B = p->d.n_problems;
  plan_bind_teq(p);
  // one cell per thread over a cluster of up to 8 CTAs of 512 threads; two cells per
  // thread beyond 4096 cells
  int cs = 1;
  while (cs < kMaxCluster && cs * 512 < Nl) cs *= 2;
  const int cpt = (Nl + cs * 512 - 1) / (cs * 512);
  int threads = (Nl + cs * cpt - 1) / (cs * cpt);
  threads = ((threads + 31) / 32) * 32;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(B * cs);
  cfg.blockDim = dim3(threads);
  cfg.stream = p->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  CK(cudaMemsetAsync(p->drec, 0, 6 * cells(p) * 8, p->stream));  // sphere_bgnd.f90:490-496
  switch (cpt) {
    case 1: CK(cudaLaunchKernelEx(&cfg, k_thin_state<1>, p->P, cs)); break;
    case 2: CK(cudaLaunchKernelEx(&cfg, k_thin_state<2>, p->P, cs)); break;
    default: return RB_ERR_UNSUPPORTED;
  }
  p->info.n_launches += 1;
  CK(cudaGetLastError());
  return RB_OK;
}

// Work list of the shared-memory column kernel: one item per (block of 32
// consecutive local cells, |mu| pair), sorted by descending ray length m+1 of the
// block's innermost cell (longest-processing-time-first).  Built from problem 0's
// edges (all problems of a plan share the shape; the order is a scheduling
// heuristic only) and cached until the edges or the cell deal change.
static int build_items(rb_plan *p, int cell_start, int cell_stride, int n_local, int nchunk) {
  const int Nl = p->d.Nl, Nmu = p->d.Nmu;
  const double *E = p->hEdges;  // solver orientation: non-increasing
  const bool far = p->use_far;
  // single-cell launches (the Gauss-Seidel order) all use the same list -- one 32-cell block,
  // every |mu| pair (x chunk): its order is only a scheduling heuristic, so it is built once
  // and kept for every cell (no host copy inside a captured sweep)
  const int cell_key = (n_local == 1) ? -2 : cell_start;
  if (p->items_start == cell_key && p->items_stride == cell_stride && p->items_far == far &&
      p->items_chunks == nchunk && p->items_edges.size() == (size_t)Nl + 1 &&
      memcmp(p->items_edges.data(), E, sizeof(double) * (Nl + 1)) == 0)
    return RB_OK;
  const int nblk = (n_local + 31) / 32, npair = (Nmu + 1) / 2;
  if (nblk > 0xfff || npair > 0xffff || nchunk > 15) return RB_ERR_UNSUPPORTED;
  const int Q = (Nl + nchunk) / nchunk;  // edges per chunk, as in sphere_ray_pair
  std::vector<std::pair<int, unsigned>> v;
  v.reserve((size_t)nblk * npair * nchunk);
  for (int ip = 0; ip < npair; ++ip) {
    const double mu = p->mu[ip];
    const double st = sqrt(1.0 - mu * mu);
    for (int jb = 0; jb < nblk; ++jb) {
      const int j = std::min(n_local - 1, jb * 32 + 31);
      const int il = cell_of(j, cell_start, cell_stride);
      const double pr = 0.5 * (E[il] + E[il + 1]) * st;
      int lo = 1, hi = Nl;  // smallest i >= 1 with E_i <= p
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (E[mid] <= pr) hi = mid; else lo = mid + 1;
      }
      const int m = lo - 1;
      int walked = m + 1;
      if (far) {   // only the edges below RATIO p are walked (sphere_ray_pair)
        const double rp = pr * RB_FAR_RATIO;
        int kf = -1;
        if (m >= 0 && E[0] >= rp) {
          int a0 = 0, a1 = m;
          while (a0 < a1) {
            const int mid = (a0 + a1 + 1) >> 1;
            if (E[mid] >= rp) a0 = mid; else a1 = mid - 1;
          }
          kf = a0;
        }
        walked = m - (kf >= 0 ? (kf & ~((1 << p->far_slog) - 1)) : -1);
      }
      for (int c = 0; c < nchunk; ++c) {  // shell edges c*Q .. (c+1)*Q-1, up to m
        const int len = (nchunk == 1) ? walked : std::min(m, (c + 1) * Q - 1) - c * Q + 1;
        // (chunks that are empty for the whole block stay in the list: their lanes store the
        // zero partials the combine kernel reads)
        v.emplace_back(-std::max(len, 0), (unsigned)jb | ((unsigned)c << 12) | ((unsigned)ip << 16));
      }
    }
  }
  std::sort(v.begin(), v.end());
  std::vector<unsigned> items(v.size());
  for (size_t i = 0; i < v.size(); ++i) items[i] = v[i].second;
  p->n_items = (int)items.size();
  CK(cudaMemcpyAsync(p->ditems, items.data(), items.size() * sizeof(unsigned),
                     cudaMemcpyHostToDevice, p->stream));
  CK(cudaStreamSynchronize(p->stream));  // `items` is pageable and dies here
  p->items_start = cell_key; p->items_stride = cell_stride; p->items_chunks = nchunk;
  p->items_far = far;
  p->items_edges.assign(E, E + Nl + 1);
  return RB_OK;
}

// opt in to more than 48 KB of shared memory per CTA (static + dynamic) where a launch needs it
template <typename K>
static int set_smem(K kern, size_t smem) {
  size_t stat = 0;
  if (smem > 0 && smem <= 48 * 1024) {
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    stat = fa.sharedSizeBytes;
  }
  if (smem + stat > 48 * 1024)
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return RB_OK;
}

// Problems of the batch that were still iterating at the last convergence read-back,
// rounded up to B / 4^k.  The launch shapes that spread ONE problem's work over the GPU
// (CTAs per problem and chunks per walk in the column kernel, cells per work unit in the
// frequency kernel, lanes per cell in the cell solver) follow it, so that the tail of a
// batch -- a handful of problems left -- still fills the SMs.  None of these choices
// changes a result bit.  Captured graphs are keyed on it (sweep_graph).
static int live_problems(const rb_plan *p) {
  const int B = p->d.n_problems;
  if (p->active_hint <= 0 || p->active_hint >= B) return B;
  int q = B;
  while (q / 4 >= p->active_hint && q / 4 >= 1) q /= 4;
  return q;
}

// far-field moment table of every problem (k_far_moments): one cluster per problem, as many
// CTAs as it takes to give every (segment, absorber) a thread and -- for small batches --
// to spread one problem's scan over several SMs
static bool far_split_wanted() {   // testing knob, part of the graph key
  const char *e = getenv("RB_FAR_SPLIT");
  return !(e && atoi(e) == 0);
}

static int launch_far_moments(rb_plan *p) {
  const int Nl = p->d.Nl, B = p->d.n_problems;
  const int nseg = (Nl + 1 + kFarSeg - 1) / kFarSeg;
  int cs = 1;
  while (cs < kMaxCluster && ((nseg + cs - 1) / cs > 42 || (long long)B * cs * 2 <= p->n_sm)) cs *= 2;
  while (cs > 1 && nseg < cs) cs /= 2;
  const int spc = (nseg + cs - 1) / cs;
  if (3 * spc > kFarThreads) return RB_ERR_UNSUPPORTED;
  // the moments of a (segment, absorber) dealt over kFarParts threads: for launches that are
  // bound by the latency of one CTA's scan (few problems), where the factor table fits in
  // shared memory (grids up to ~6500 shells).  Large batches keep one thread per (segment,
  // absorber): six CTAs per SM instead of one.  RB_FAR_SPLIT=0: never split.
  const bool split = far_split_wanted() && (long long)B * cs <= 2LL * p->n_sm &&
                     far_split_smem_bytes(spc) <= (size_t)226 * 1024;
  const size_t smem = split ? far_split_smem_bytes(spc) : far_smem_bytes(spc);
  int rc = split ? set_smem(k_far_moments_split, smem) : set_smem(k_far_moments, smem);
  if (rc) return rc;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)B * cs);
  cfg.blockDim = dim3(kFarThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = p->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (split) CK(cudaLaunchKernelEx(&cfg, k_far_moments_split, p->P, p->dfar, cs, spc, p->far_slog));
  else CK(cudaLaunchKernelEx(&cfg, k_far_moments, p->P, p->dfar, cs, spc, p->far_slog));
  p->info.n_launches += 1;
  return RB_OK;
}

static int launch_sphere_columns(rb_plan *p, int cell_start, int cell_stride, int n_local) {
  const int Nl = p->d.Nl, B = p->d.n_problems;
  cudaStream_t s = p->stream;
  const size_t rec_smem = (size_t)(Nl + 1 + kPackPad) * sizeof(double4);
  const size_t smem = p->use_far ? far_cols_smem(Nl, p->far_slog) : rec_smem;
  const size_t kMaxSmem = 227 * 1024 - 1024;
  if (smem > kMaxSmem) {  // records do not fit in shared memory: global-memory fallback
    dim3 gp((Nl + 1 + kPackPad + 255) / 256, B);
    k_pack_shells<<<gp, 256, 0, s>>>(p->P);
    dim3 gc((n_local + 127) / 128, (p->d.Nmu + 1) / 2, B);
    k_sphere_columns<128, 1, 6><<<gc, 128, 0, s>>>(p->P, cell_start, cell_stride, n_local);
    p->info.n_launches += 2;
    return RB_OK;
  }
  // as many CTAs per SM as the records allow (up to 3), 16-18 warps per SM in all
  int cps = (int)std::min<size_t>(3, kMaxSmem / (smem + 1024));
  const int warps = (cps == 1) ? 16 : (cps == 2 ? 8 : 6);
  const int want = p->n_sm * cps;  // CTAs that fill the GPU once
  // too few rays per launch to keep every warp busy with whole walks (cell shards over many
  // GPUs, small grids): cut every walk into nchunk pieces (separate work items)
  int nchunk = 1;
  {
    const long long lane_items =
        (long long)((n_local + 31) / 32) * ((p->d.Nmu + 1) / 2) * live_problems(p);
    const long long warps_total = (long long)p->n_sm * warps * cps;
    while (nchunk < 8 && lane_items * nchunk < warps_total * 6) nchunk *= 2;  // >= 6 items per warp
    const char *cenv = getenv("RB_COLS_CHUNKS");  // testing knob
    if (cenv && atoi(cenv) >= 1 && atoi(cenv) <= 8) nchunk = atoi(cenv);
    while (nchunk > 1 && Nl < 8 * nchunk) nchunk /= 2;
    // the chunk partials are indexed by problem: keep them under 2 GiB for any batch size
    // (the tail of a large batch then walks unchunked)
    auto part_bytes = [&](int nc) {
      return (size_t)B * ((p->d.Nmu + 1) / 2) * n_local * nc * kPartStride * sizeof(double);
    };
    while (nchunk > 1 && part_bytes(nchunk) > ((size_t)2 << 30)) nchunk /= 2;
    if (p->use_far) nchunk = 1;   // the walks are short: whole items balance well enough
  }
  int rc = build_items(p, cell_start, cell_stride, n_local, nchunk);
  if (rc) return rc;
  if (nchunk > 1) {
    const size_t need = (size_t)B * ((p->d.Nmu + 1) / 2) * n_local * nchunk * kPartStride * 8;
    if (need > p->part_bytes) {
      cudaFree(p->dpart);
      p->dpart = nullptr; p->part_bytes = 0;
      if (cudaMalloc((void **)&p->dpart, need) != cudaSuccess) {
        cudaGetLastError();
        return RB_ERR_NOMEM;
      }
      p->part_bytes = need;
    }
  }

  if (p->use_far) {
    rc = launch_far_moments(p);
    if (rc) return rc;
  }
  const int waves = (B > 1) ? 4 : 1;  // batches: ~4 CTAs per resident slot (finer tail)
  const int live = live_problems(p);
  // Launch shape where the records allow one CTA per SM (the same operations in the same
  // order in every shape: results are bit-identical).  A full launch (C4 on one GPU: 110 ray-pair
  // items per SM) is throughput-bound and runs 16 warps per SM.  Cell shards have fewer, and
  // the launch is then bound by the latency of the warps that hold the long walks: fewer warps
  // per SM run them faster -- one rank's share of the C4 column stage on 2 / 4 / 8 GPUs:
  // 0.130 / 0.107 / 0.096 ms with 512 threads, 0.125 / 0.093 / 0.084 with 384, 0.138 / 0.090 /
  // 0.077 with 256 (profiles/r02_kernel_variants.md).  RB_COLS_VARIANT (testing knob): 1 = 256
  // threads x 4 roots per trip, 6 = 384 x 1, 7 = 512 x 1, 9 = 256 x 1.
  int variant = 7;
  if (p->use_far && nchunk == 1 && cps == 1) {
    const long long items = (long long)p->n_items * live, full = (long long)p->n_sm * 16;
    if (2 * items < 5 * full) variant = 9;
    else if (items < 5 * full) variant = 6;
  }
  if (const char *se = getenv("RB_COLS_VARIANT")) {
    const int v = atoi(se);
    if (v == 1 || v == 6 || v == 7 || v == 9) variant = v;
  }
  const int warps_cta = cps != 1 ? warps : (variant == 7 ? 16 : (variant == 6 ? 12 : 8));
  int per_problem = std::max(1, (waves * want + live - 1) / live);
  per_problem = std::min(per_problem, (p->n_items + warps_cta - 1) / warps_cta);
  // the grid spans all B problems (converged ones return at once): bound it
  per_problem = std::min(per_problem, std::max(1, (1 << 17) / B));
  dim3 g(per_problem, B);
  CK(cudaMemsetAsync(p->dwork, 0, B * sizeof(int), s));
#define COLS(THREADS, U, MINB)                                                               \
  do {                                                                                       \
    rc = set_smem(k_sphere_columns_smem<THREADS, U, MINB>, smem);                            \
    if (rc) return rc;                                                                       \
    k_sphere_columns_smem<THREADS, U, MINB><<<g, THREADS, smem, s>>>(                        \
        p->P, cell_start, cell_stride, n_local, p->ditems, p->n_items, p->dwork, nchunk,     \
        p->dpart, p->use_far ? p->dfar : nullptr, p->far_slog);                              \
  } while (0)
  // 16-18 warps per SM at ~100 registers beat 24 warps at 80 (1.035 vs 1.068 ms at C4,
  // profiles/r01_kernel_variants.md)
  if (cps == 1) {
    switch (variant) {
      case 1: COLS(256, 4, 1); break;
      case 6: COLS(384, 1, 1); break;
      case 9: COLS(256, 1, 1); break;
      default: COLS(512, 1, 1); break;
    }
  }
  else if (cps == 2) COLS(256, 1, 2);
  else COLS(192, 1, 3);
#undef COLS
  p->info.n_launches += 1;
  if (nchunk > 1) {
    dim3 gm((n_local + 127) / 128, (p->d.Nmu + 1) / 2, B);
    k_columns_combine<<<gm, 128, 0, s>>>(p->P, cell_start, cell_stride, n_local, nchunk,
                                         p->dpart);
    p->info.n_launches += 1;
  }
  return RB_OK;
}

static int launch_nu_rates(rb_plan *p, int cell_start, int cell_stride, int n_local) {
  const int B = p->d.n_problems, Nnu = p->d.Nnu, g = p->d.geometry;
  cudaStream_t s = p->stream;
  int rc = RB_OK;
#define NU(ATT, KU, RU, BLOCK, MINB, LTAB, CPB)                                              \
  do {                                                                                       \
    const int ndir_pad = (p->Ndir + RU - 1) / RU * RU;                                       \
    const size_t smem = (size_t)ndir_pad * sizeof(double4) * CPB;                            \
    rc = set_smem(k_nu_rates<ATT, KU, RU, BLOCK, MINB, LTAB, CPB>, smem);                    \
    if (rc) return rc;                                                                       \
    int occ = 1;                                                                             \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(                                           \
        &occ, k_nu_rates<ATT, KU, RU, BLOCK, MINB, LTAB, CPB>, BLOCK, smem);                 \
    /* 4 CTAs per resident slot: the hardware scheduler evens out the tail (+3 %) */       \
    const long long want = (long long)p->n_sm * std::max(1, occ) * 4;                        \
    /* cells per work unit: 16 in a full batch (a CTA re-reads one problem's frequency   */ \
    /* table from L1), fewer when few problems are left (>= 2 units per resident cell    */ \
    /* slot), bounded below so that the flat unit list stays under 2^20 entries          */ \
    int cpu = 1;                                                                             \
    if (B > 1) {                                                                             \
      const long long slots = (long long)p->n_sm * std::max(1, occ) * CPB;                   \
      const long long fit = (long long)live_problems(p) * n_local / (2 * slots);             \
      cpu = (int)std::min<long long>(16, std::max<long long>(1, fit));                       \
      while (cpu < 16 && (long long)B * ((n_local + cpu - 1) / cpu) > (1ll << 20)) cpu *= 2; \
      cpu = std::min(cpu, n_local);                                                          \
    }                                                                                        \
    const long long units = (long long)((n_local + cpu - 1) / cpu) * B;                      \
    dim3 gn((unsigned)std::min<long long>(want, (units + CPB - 1) / CPB));                   \
    k_nu_rates<ATT, KU, RU, BLOCK, MINB, LTAB, CPB><<<gn, BLOCK, smem, s>>>(                 \
        p->P, cell_start, cell_stride, n_local, cpu);                                        \
  } while (0)
  // thread <-> 4 frequencies, 2 rays in flight: 8 independent exp chains per thread
  if (g == RB_GEOM_SLAB_BGND) {
    if (Nnu <= 32) NU(1, 1, 1, 32, 8, 0, 1);
    else if (Nnu <= 64) NU(1, 1, 1, 64, 4, 0, 1);
    else NU(1, 1, 1, 128, 3, 0, 1);
  } else {
    // timed on B200 (profiles/r01_kernel_variants.md, r02_kernel_variants.md): 4 frequencies
    // x 2 rays in flight for wide grids; one warp per cell, 4 frequencies x 1 ray, 16 warps
    // per SM for Nnu <= 128.  exp table: 1024 entries x 4 copies + cubic (11 FP64 per
    // evaluation) where 128 threads share it -- one cell per CTA on wide grids, four cells
    // (one per warp) on narrow ones; RB_NU_LTAB=8 / 6 select the round-1 tables (256
    // entries, degree 4 / 64 entries, degree 5, one 32-thread CTA per cell) for A/B runs
    const char *le = getenv("RB_NU_LTAB");
    const int lt = le ? atoi(le) : 10;
    if (Nnu <= 128) { if (lt == 10) NU(0, 4, 1, 128, 4, 10, 4); else NU(0, 4, 1, 32, 16, 6, 1); }
    else if (Nnu <= 256) NU(0, 4, 2, 64, 6, 6, 1);
    else { if (lt == 10) NU(0, 4, 2, 128, 4, 10, 1); else NU(0, 4, 2, 128, 4, 8, 1); }
  }
#undef NU
  p->info.n_launches += 1;
  return rc;
}

// isotropic photon transfer: work list (block of 32 cells in the routine's outermost-first
// order, ray), sorted by descending path length of the block's innermost cell
// (mu <= 0: 2m - il + 1 shells, mu > 0: il + 1); cached until the edges change
static int build_iso_items(rb_plan *p, const RecCells &L) {
  const int Nl = p->d.Nl, Nmu = p->d.Nmu;
  const double *E = p->hEdges;  // problem 0, plan orientation
  if (p->iso_edges.size() == (size_t)Nl + 1 && p->iso_cells[0] == L.start &&
      p->iso_cells[1] == L.stride && p->iso_cells[2] == L.n_total &&
      memcmp(p->iso_edges.data(), E, sizeof(double) * (Nl + 1)) == 0)
    return RB_OK;
  const bool of = p->d.geometry == RB_GEOM_SPHERE_BGND;  // outermost-first storage
  auto Eo = [&](int i) { return of ? E[i] : E[Nl - i]; };   // edge i, outermost first
  const int nblk = (L.n_total + 31) / 32;
  std::vector<std::pair<int, unsigned>> v;
  v.reserve((size_t)nblk * Nmu);
  for (int imu = 0; imu < Nmu; ++imu) {
    const double mu = p->mu[imu];
    const double st = sqrt(1.0 - mu * mu);
    for (int jb = 0; jb < nblk; ++jb) {
      int len = 0;   // the longer of the block's two end cells
      for (int e = 0; e < 2; ++e) {
        const int j = e ? std::min(L.n_total - 1, jb * 32 + 31) : jb * 32;
        const int pc = L.start + j * L.stride;
        const int il = of ? pc : Nl - 1 - pc;
        const double pr = 0.5 * (Eo(il) + Eo(il + 1)) * st;
        int lo = 1, hi = Nl;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (Eo(mid) <= pr) hi = mid; else lo = mid + 1;
        }
        const int m = lo - 1;
        len = std::max(len, (mu <= 0.0) ? 2 * m - il + 1 : il + 1);
      }
      v.emplace_back(-len, (unsigned)jb | ((unsigned)imu << 16));
    }
  }
  std::sort(v.begin(), v.end());
  std::vector<unsigned> items(v.size());
  for (size_t i = 0; i < v.size(); ++i) items[i] = v[i].second;
  p->n_iso_items = (int)items.size();
  CK(cudaMemcpyAsync(p->diso_items, items.data(), items.size() * sizeof(unsigned),
                     cudaMemcpyHostToDevice, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  p->iso_edges.assign(E, E + Nl + 1);
  p->iso_cells[0] = L.start; p->iso_cells[1] = L.stride; p->iso_cells[2] = L.n_total;
  return RB_OK;
}

// recombination-photon rates of the cells of list L from the current state (rb_rec.cuh);
// the per-cell preparation (emissivities, shell optical depths) covers all cells
static int launch_rec(rb_plan *p, const RecCells &L) {
  const int Nl = p->d.Nl, B = p->d.n_problems, g = p->d.geometry;
  const int nc = L.n_total;
  cudaStream_t s = p->stream;
  const bool sphere = g == RB_GEOM_SPHERE_BGND || g == RB_GEOM_SPHERE_STROMGREN;
  dim3 gp((Nl + 1 + 127) / 128, B);
  k_rec_prep<<<gp, 128, 0, s>>>(p->P, p->R, p->K, p->rec_meth);
  dim3 gc((nc + 63) / 64, B);
  if (nc == 0) {
    p->info.n_launches += 1;
    return RB_OK;
  }
  if (!sphere) {
    k_rec_slab_prefix<<<B, 32, 0, s>>>(p->P, p->R, p->K);
    // J of the three lines goes through the (unused for slabs) F0 scratch: [B][3][Nl]
    double *J = p->R.F0[0];
    dim3 gw((nc + 3) / 4, 3, B);   // one warp per (layer, line)
    k_rec_slab<<<gw, 128, 0, s>>>(p->P, p->R, J, L);
    k_rec_slab_finish<<<gc, 64, 0, s>>>(p->P, p->R, p->K, J, L);
    p->info.n_launches += 4;
  } else if (p->rec_meth == 2) {
    dim3 gw((nc + 3) / 4, B);   // one warp per solve layer
    k_rec_outward<<<gw, 128, 0, s>>>(p->P, p->R, p->K, L);
    p->info.n_launches += 2;
  } else if (p->rec_meth == 3) {
    dim3 gw((nc + 3) / 4, B);
    k_rec_radial<<<gw, 128, 0, s>>>(p->P, p->R, p->K, L);
    p->info.n_launches += 2;
  } else {
    const size_t smem = (size_t)(Nl + 1) * 56;
    if (smem <= 227 * 1024 - 1024 && (Nl + 31) / 32 <= 0xffff) {
      int rc = build_iso_items(p, L);
      if (rc) return rc;
      CK(cudaMemsetAsync(p->dwork, 0, B * sizeof(int), s));
      rc = set_smem(k_rec_iso_rays_smem<512>, smem);
      if (rc) return rc;
      const int per_problem = std::max(1, std::min((p->n_iso_items + 15) / 16,
                                                   ((B > 1 ? 4 : 1) * p->n_sm + B - 1) / B));
      dim3 gr(per_problem, B);
      k_rec_iso_rays_smem<512><<<gr, 512, smem, s>>>(p->P, p->R, p->d.Nmu, p->diso_items,
                                                      p->n_iso_items, p->dwork, L);
    } else {
      dim3 gr((nc + 127) / 128, p->d.Nmu, B);
      k_rec_iso_rays<<<gr, 128, 0, s>>>(p->P, p->R, p->d.Nmu, L);
    }
    k_rec_iso_finish<<<gc, 64, 0, s>>>(p->P, p->R, p->K, p->d.Nmu, L);
    p->info.n_launches += 3;
  }
  CK(cudaGetLastError());
  return RB_OK;
}

// Lanes per cell in the cell solver.  G lanes run the bisection (on T with find_Teq:
// k_cell_solve_spec, on n_e at fixed T: k_cell_solve_pce_spec) speculatively; G is the
// largest group that keeps the cells of the problems STILL ITERATING (as of the last
// convergence read-back: the tail of a batch is a few problems, whose serial bisections
// would be pure latency) within ~one wave of resident threads and the grid - which spans
// all problems, converged ones return at once - below 2^17 CTAs; larger batches run one
// thread per cell.  Every G gives the serial algorithm's results bit for bit.
static int cell_group_size(const rb_plan *p, int n_local) {
  const int B = p->d.n_problems;
  const char *fenv = getenv("RB_SPEC_G");  // testing knob: force the group size
  const int forced = fenv ? atoi(fenv) : 0;
  if (forced == 1 || forced == 4 || forced == 8 || forced == 16 || forced == 32) return forced;
  const long long cells_now = (long long)n_local * live_problems(p);
  const long long budget = (long long)p->n_sm * 512;
  int G = 32;
  while (G > 1 && cells_now * G > budget) G >>= 1;
  while (G > 1 && ((long long)n_local * G + 127) / 128 * B > (1ll << 17)) G >>= 1;
  return G < 4 ? 1 : G;
}

// G lanes per cell run the bisection (on T with find_Teq: k_cell_solve_spec, on n_e at
// fixed T: k_cell_solve_pce_spec) speculatively; G is the largest group that keeps the
// launch within ~one wave of resident threads, larger batches run one thread per cell.
static void launch_cell_solve(rb_plan *p, const PeerDev &Q, int cell_start, int cell_stride,
                              int n_local) {
  const int B = p->d.n_problems;
  cudaStream_t s = p->stream;
  const int G = cell_group_size(p, n_local);
  if (G == 1) {
    dim3 gs3((n_local + 127) / 128, B);
    k_cell_solve<<<gs3, 128, 0, s>>>(p->P, Q, cell_start, cell_stride, n_local);
  } else {
    dim3 gs3(((long long)n_local * G + 127) / 128, B);
#define CELL(KERN)                                                                           \
  switch (G) {                                                                               \
    case 32: KERN<32><<<gs3, 128, 0, s>>>(p->P, Q, cell_start, cell_stride, n_local); break; \
    case 16: KERN<16><<<gs3, 128, 0, s>>>(p->P, Q, cell_start, cell_stride, n_local); break; \
    case 8: KERN<8><<<gs3, 128, 0, s>>>(p->P, Q, cell_start, cell_stride, n_local); break;   \
    default: KERN<4><<<gs3, 128, 0, s>>>(p->P, Q, cell_start, cell_stride, n_local);         \
  }
    if (p->d.i_find_Teq) { CELL(k_cell_solve_spec) } else { CELL(k_cell_solve_pce_spec) }
#undef CELL
  }
}

// One sweep of SphereBgnd in the reference's sequential order (sphere_bgnd.f90:753-913 with
// OMP_NUM_THREADS=1): the recombination-photon rates once from the pre-sweep state
// (:694-735), then the cells outermost first, each one seeing the fractions and temperatures
// of the cells before it ALREADY UPDATED -- Gauss-Seidel.  The three stages run per cell
// (columns of its rays from the current state -> frequency / angle integrals -> cell solve):
// 4-5 small launches per cell, strictly serial.  This is the reference's only deterministic
// order; it is a non-default switch because it cannot use the GPU (or shard), see DESIGN.md.
static int sweep_gauss_seidel(rb_plan *p) {
  const int Nl = p->d.Nl;
  cudaStream_t s = p->stream;
  cudaEvent_t *sev = p->sev[p->sweeps_issued % kSlots];
  const bool timed = !p->capturing;
  p->slot_timed[p->sweeps_issued % kSlots] = timed;
  if (timed) cudaEventRecord(p->rev[p->sweeps_issued % kSlots], s);
  int rc = RB_OK;
  if (p->rec_meth != 1) {
    RecCells L = {0, 1, Nl, Nl};
    rc = launch_rec(p, L);
    if (rc) return rc;
  }
  if (timed) cudaEventRecord(sev[0], s);
  PeerDev Q = p->Q;
  Q.world = 1;
  // one cell per launch: the far-field table would be rebuilt for every cell -- walk instead
  const bool far_saved = p->use_far;
  p->use_far = false;
  for (int il = 0; il < Nl && !rc; ++il) {
    rc = launch_sphere_columns(p, il, Nl, 1);
    if (!rc) rc = launch_nu_rates(p, il, Nl, 1);
    if (!rc) {
      launch_cell_solve(p, Q, il, Nl, 1);
      p->info.n_launches += 1;
    }
  }
  p->use_far = far_saved;
  if (rc) return rc;
  if (timed) {   // the stages interleave: the whole cell loop is booked as "columns"
    cudaEventRecord(sev[1], s); cudaEventRecord(sev[2], s); cudaEventRecord(sev[3], s);
    CK(cudaGetLastError());
  }
  return RB_OK;
}

static int sweep_impl(rb_plan *p, int cell_start, int cell_stride, bool scatter) {
  if (!p) return RB_ERR_BAD_ARG;
  PLAN_DEV(p);
  if (p->sweep_order == RB_ORDER_GAUSS_SEIDEL && p->d.geometry == RB_GEOM_SPHERE_BGND) {
    // sequential in the cells: no cell shards (replicas only in this mode)
    if (scatter || cell_start != 0 || cell_stride != 1) return RB_ERR_UNSUPPORTED;
    return sweep_gauss_seidel(p);
  }
  PeerDev Q = p->Q;
  if (!scatter) Q.world = 1;
  const int Nl = p->d.Nl, B = p->d.n_problems, g = p->d.geometry;
  const bool two = (g == RB_GEOM_SLAB_PLANE_2 || g == RB_GEOM_SLAB_BGND);
  const int ncell = two ? Nl / 2 : Nl;  // slab_bgnd.f90:686 (Nl2 = Nl/2)
  const int n_local = deal_n_local(ncell, cell_start, cell_stride);
  if (n_local < 0) return RB_ERR_BAD_ARG;
  cudaStream_t s = p->stream;
  cudaEvent_t *sev = p->sev[p->sweeps_issued % kSlots];
  int rc = RB_OK;
  const bool timed = !p->capturing;   // stage events are not captured into graphs
  p->slot_timed[p->sweeps_issued % kSlots] = timed;
  if (timed) cudaEventRecord(p->rev[p->sweeps_issued % kSlots], s);
  if (p->rec_meth != 1) {
    // photon transfer for this rank's solve layers.  One GPU: all Nl layers (slabs too:
    // the reference fills both halves before the mirrored copy, slab_base.f90:163).
    RecCells L = {cell_start, cell_stride, n_local, n_local};
    if (two) {
      if (cell_start == 0 && cell_stride == 1) L.n_local = L.n_total = Nl;
      else L.n_total = n_local + (Nl & 1);   // + the never-solved middle layer of an odd Nl
    }
    rc = launch_rec(p, L);
    if (rc) return rc;
  }
  if (timed) cudaEventRecord(sev[0], s);
  if (g == RB_GEOM_SPHERE_BGND) {
    if (n_local > 0) rc = launch_sphere_columns(p, cell_start, cell_stride, n_local);
    if (rc) return rc;
  } else {
    k_column_taus<512><<<B, 512, 0, s>>>(p->P);
    p->info.n_launches += 1;
  }
  if (timed) cudaEventRecord(sev[1], s);
  if (n_local > 0) {
    rc = launch_nu_rates(p, cell_start, cell_stride, n_local);
    if (rc) return rc;
    if (timed) cudaEventRecord(sev[2], s);
    launch_cell_solve(p, Q, cell_start, cell_stride, n_local);
    p->info.n_launches += 1;
  } else {
    if (timed) cudaEventRecord(sev[2], s);
  }
  if (timed) cudaEventRecord(sev[3], s);
  if (timed) CK(cudaGetLastError());
  return RB_OK;
}

extern "C" int rb_plan_set_sweep_order(rb_plan *p, int order) {
  if (!p || (order != RB_ORDER_JACOBI && order != RB_ORDER_GAUSS_SEIDEL)) return RB_ERR_BAD_ARG;
  if (order == RB_ORDER_GAUSS_SEIDEL && p->d.geometry != RB_GEOM_SPHERE_BGND)
    return RB_ERR_UNSUPPORTED;   // the other three drivers ARE Jacobi in the reference
  p->sweep_order = order;        // part of the graph key
  return RB_OK;
}

extern "C" int rb_plan_sweep_local(rb_plan *p, int cell_start, int cell_stride) {
  return sweep_impl(p, cell_start, cell_stride, false);
}

// ---------------------------------------------------------------------------
// cell sharding over peer memory
// ---------------------------------------------------------------------------
static int plan_sweep_cells(const rb_plan *p);

static unsigned long long peer_timeout_ns() {
  // RB_PEER_TIMEOUT_MS: how long a rank waits for its peers at the per-sweep barrier
  const char *e = getenv("RB_PEER_TIMEOUT_MS");
  double ms = e ? atof(e) : 0.0;
  if (!(ms > 0.0)) ms = 2000.0;
  return (unsigned long long)(ms * 1.0e6);
}

extern "C" int rb_plan_peer_alloc(rb_plan *p, int rank, int world, unsigned char handle[64]) {
  if (!p || !handle || world < 1 || world > kMaxPeers || rank < 0 || rank >= world ||
      p->d.n_problems != 1 || p->dpeer)
    return RB_ERR_BAD_ARG;
  PLAN_DEV(p);
  const int ncell = plan_sweep_cells(p);
  if (ncell % world != 0) return RB_ERR_BAD_ARG;
  const int n_local = ncell / world;
  const size_t bytes = 256 + (size_t)2 * world * kPeerFields * n_local * sizeof(double);
  if (cudaMalloc(&p->dpeer, bytes) != cudaSuccess) {
    cudaGetLastError();
    p->dpeer = nullptr;
    return RB_ERR_NOMEM;
  }
  CK(cudaMemset(p->dpeer, 0, bytes));
  CK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, p->dpeer));
  static_assert(sizeof(h) == 64, "CUDA IPC handle size");
  memcpy(handle, &h, 64);
  p->Q.world = 1;  // not usable until rb_plan_peer_connect
  p->Q.rank = rank; p->Q.n_local = n_local;
  // 32-cell blocks dealt over the ranks where the grid allows it (compact ray walks per
  // warp, see cell_of()); RB_PEER_DEAL=cyclic keeps single cells
  const char *de = getenv("RB_PEER_DEAL");
  const bool cyc = (de && (de[0] == 'c' || de[0] == 'C')) || ncell % (32 * world) != 0;
  p->Q.deal = cyc ? world : -world;
  p->peer_world = world;
  p->peer_base[rank] = p->dpeer;
  return RB_OK;
}

extern "C" int rb_plan_peer_connect(rb_plan *p, const unsigned char *handles) {
  if (!p || !handles || !p->dpeer) return RB_ERR_BAD_ARG;
  PLAN_DEV(p);
  const int world = p->peer_world, rank = p->Q.rank;
  for (int r = 0; r < world; ++r) {
    if (r == rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * 64, 64);
    CK(cudaIpcOpenMemHandle(&p->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
  }
  for (int r = 0; r < world; ++r) {
    p->Q.flags[r] = (unsigned long long *)p->peer_base[r];
    p->Q.stage[r] = (double *)((char *)p->peer_base[r] + 256);
  }
  p->Q.ctl = (unsigned long long *)p->dpeer + kPeerCtl;
  p->Q.timeout_ns = peer_timeout_ns();
  p->Q.world = world;
  p->peer_sweeps = 0;
  destroy_graphs(p);
  return RB_OK;
}

// the barrier + scatter + convergence test that ends a cell-sharded sweep, and its read-back
static int exchange_body(rb_plan *p, const PeerDev &Q, int slot) {
  cudaStream_t s = p->stream;
  const int ncell = Q.world * Q.n_local;
  const int ctas = std::max(1, std::min(32, (ncell + 255) / 256));
  k_exchange_finish<256><<<ctas, 256, 0, s>>>(p->P, Q, slot);
  p->info.n_launches += 1;
  CK(cudaMemcpyAsync(p->hflags + 3 + slot, p->dflags + 3 + slot, sizeof(int),
                     cudaMemcpyDeviceToHost, s));
  return RB_OK;
}

extern "C" int rb_plan_sweep_sharded(rb_plan *p) {
  if (!p || p->Q.world < 2) return RB_ERR_BAD_ARG;
  if (p->sweeps_issued - p->sweeps_polled >= kSlots - 1) return RB_ERR_BAD_ARG;  // poll first
  PLAN_DEV(p);
  int rc = sweep_impl(p, p->Q.rank, p->Q.deal, true);
  if (rc) return rc;
  const int slot = (int)(p->sweeps_issued % kSlots);
  PeerDev Q = p->Q;
  // the first barrier after connecting also absorbs the peers' one-off costs (module load,
  // work-list build with its host synchronisation): ten times the timeout
  if (p->peer_sweeps == 0) Q.timeout_ns *= 10ull;
  rc = exchange_body(p, Q, slot);
  if (rc) return rc;
  CK(cudaEventRecord(p->sev[slot][4], p->stream));
  p->sweeps_issued += 1;
  p->peer_sweeps += 1;
  CK(cudaGetLastError());
  return RB_OK;
}

// convergence test + bookkeeping of the sweep just enqueued; does not wait.
static int finish_body(rb_plan *p, int slot) {
  const int B = p->d.n_problems;
  cudaStream_t s = p->stream;
  k_converge_finish<1024><<<B, 1024, 0, s>>>(p->P, slot);
  p->info.n_launches += 1;
  CK(cudaMemcpyAsync(p->hflags + 3 * B + slot, p->dflags + 3 * B + slot, sizeof(int),
                     cudaMemcpyDeviceToHost, s));
  return RB_OK;
}

extern "C" int rb_plan_finish_sweep_async(rb_plan *p) {
  if (!p) return RB_ERR_BAD_ARG;
  PLAN_DEV(p);
  if (p->sweeps_issued - p->sweeps_polled >= kSlots - 1) return RB_ERR_BAD_ARG;  // poll first
  const int slot = (int)(p->sweeps_issued % kSlots);
  int rc = finish_body(p, slot);
  if (rc) return rc;
  CK(cudaEventRecord(p->sev[slot][4], p->stream));
  p->sweeps_issued += 1;
  CK(cudaGetLastError());
  return RB_OK;
}

// One whole sweep (all cells) + convergence test as a CUDA graph launch.  A sweep is 4-7
// small launches; for the column geometries and small grids they cost more than the kernels
// (70-100 us per sweep of ~40 us of work), so rb_plan_solve replays a captured graph per
// in-flight slot (the slot index is a kernel argument).  Graphs are re-captured when a
// kernel argument changed (the plan was re-armed with other flags).  Stage events are not
// part of the graphs: per-stage times come from direct launches (first sweep, bench).
static int sweep_graph(rb_plan *p, bool sharded) {
  if (p->sweeps_issued - p->sweeps_polled >= kSlots - 1) return RB_ERR_BAD_ARG;
  const int slot = (int)(p->sweeps_issued % kSlots);
  cudaStream_t s = p->stream;
  rb_plan::GraphKey key;
  memset(&key, 0, sizeof(key));
  memcpy(&key.P, &p->P, sizeof(PlanDev));
  memcpy(&key.K, &p->K, sizeof(RecConst));
  memcpy(&key.R, &p->R, sizeof(RecDev));
  if (sharded) memcpy(&key.Q, &p->Q, sizeof(PeerDev));
  key.rec_meth = p->rec_meth;
  key.live = live_problems(p);
  key.sharded = sharded ? 1 : 0;
  key.use_far = p->use_far ? 1 : 0;
  key.order = p->sweep_order;
  {
    const char *a = getenv("RB_SPEC_G"), *b = getenv("RB_COLS_CHUNKS"), *c = getenv("RB_NU_LTAB");
    key.spec_g = a ? atoi(a) : 0;
    key.cols_chunks = b ? atoi(b) : 0;
    key.nu_ltab = c ? atoi(c) : 0;
    key.far_split = far_split_wanted() ? 1 : 0;
    const char *cv = getenv("RB_COLS_VARIANT");
    key.cols_variant = cv ? atoi(cv) : 0;
  }
  if (!p->gkey_valid || memcmp(&p->gkey, &key, sizeof(key)) != 0) {
    destroy_graphs(p);
    memcpy(&p->gkey, &key, sizeof(key));
    p->gkey_valid = true;
  }
  if (!p->gexec[slot]) {
    const int nl0 = p->info.n_launches;
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      p->graphs_ok = false;
      return RB_ERR_UNSUPPORTED;
    }
    p->capturing = true;
    int rc;
    if (sharded) {
      rc = sweep_impl(p, p->Q.rank, p->Q.deal, true);
      if (!rc) rc = exchange_body(p, p->Q, slot);
    } else {
      rc = sweep_impl(p, 0, 1, false);
      if (!rc) rc = finish_body(p, slot);
    }
    p->capturing = false;
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(s, &graph);
    p->graph_nl[slot] = p->info.n_launches - nl0;
    p->info.n_launches = nl0;
    if (rc || e != cudaSuccess || !graph ||
        cudaGraphInstantiate(&p->gexec[slot], graph, 0) != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      p->gexec[slot] = nullptr;
      p->graphs_ok = false;
      return rc ? rc : RB_ERR_UNSUPPORTED;
    }
    cudaGraphDestroy(graph);
  }
  p->slot_timed[slot] = false;
  CK(cudaGraphLaunch(p->gexec[slot], s));
  p->info.n_launches += p->graph_nl[slot];
  CK(cudaEventRecord(p->sev[slot][4], s));
  p->sweeps_issued += 1;
  if (sharded) p->peer_sweeps += 1;
  return RB_OK;
}

// wait for the oldest un-polled sweep; returns the number of problems still active after it
extern "C" int rb_plan_poll_sweep(rb_plan *p, int *n_active) {
  if (!p || p->sweeps_polled >= p->sweeps_issued) return RB_ERR_BAD_ARG;
  PLAN_DEV(p);
  const int B = p->d.n_problems;
  const int slot = (int)(p->sweeps_polled % kSlots);
  CK(cudaEventSynchronize(p->sev[slot][4]));
  float ms;  // per-stage device times of that sweep
  cudaEvent_t *sev = p->sev[slot];
  if (p->slot_timed[slot]) {
    if (cudaEventElapsedTime(&ms, p->rev[slot], sev[0]) == cudaSuccess) p->info.ms_rec += ms;
    if (cudaEventElapsedTime(&ms, sev[0], sev[1]) == cudaSuccess) p->info.ms_columns += ms;
    if (cudaEventElapsedTime(&ms, sev[1], sev[2]) == cudaSuccess) p->info.ms_nu += ms;
    if (cudaEventElapsedTime(&ms, sev[2], sev[3]) == cudaSuccess) p->info.ms_cell += ms;
    if (cudaEventElapsedTime(&ms, sev[3], sev[4]) == cudaSuccess) p->info.ms_finish += ms;
    cudaGetLastError();
  }
  p->sweeps_polled += 1;
  p->info.iters += 1;
  p->last_active = p->hflags[3 * B + slot];
  if (n_active) *n_active = p->last_active;
  return RB_OK;
}

extern "C" int rb_plan_finish_sweep(rb_plan *p, int *n_active) {
  int rc = rb_plan_finish_sweep_async(p);
  if (rc) return rc;
  int n = 0;
  while (p->sweeps_polled < p->sweeps_issued) {
    rc = rb_plan_poll_sweep(p, &n);
    if (rc) return rc;
  }
  if (n_active) *n_active = n;
  return RB_OK;
}

static int plan_sweep_cells(const rb_plan *p) {
  const int g = p->d.geometry;
  const bool two = (g == RB_GEOM_SLAB_PLANE_2 || g == RB_GEOM_SLAB_BGND);
  return two ? p->d.Nl / 2 : p->d.Nl;
}

extern "C" int rb_plan_pack_cells(rb_plan *p, int cell_start, int cell_stride, double *dev_send) {
  if (!p || !dev_send || p->d.n_problems != 1) return RB_ERR_BAD_ARG;
  PLAN_DEV(p);
  const int ncell = plan_sweep_cells(p);
  const int n_local = deal_n_local(ncell, cell_start, cell_stride);
  if (n_local < 0) return RB_ERR_BAD_ARG;
  if (n_local > 0) {
    dim3 g((n_local + 127) / 128, kExchangeFields);
    k_pack_cells<<<g, 128, 0, p->stream>>>(p->P, cell_start, cell_stride, n_local, dev_send);
    p->info.n_launches += 1;
  }
  CK(cudaGetLastError());
  return RB_OK;
}

extern "C" int rb_plan_unpack_cells(rb_plan *p, int world, const double *dev_recv) {
  // world > 0: cyclic deal il = j*world + g; world < 0: 32-cell blocks over -world ranks
  if (!p || !dev_recv || world == 0 || p->d.n_problems != 1) return RB_ERR_BAD_ARG;
  PLAN_DEV(p);
  const int ncell = plan_sweep_cells(p);
  const int W = world > 0 ? world : -world;
  if (ncell % W != 0 || deal_n_local(ncell, 0, world) < 0) return RB_ERR_BAD_ARG;
  const int n_local = ncell / W;
  dim3 g((n_local + 127) / 128, kExchangeFields, W);
  k_unpack_cells<<<g, 128, 0, p->stream>>>(p->P, world, n_local, dev_recv);
  p->info.n_launches += 1;
  CK(cudaGetLastError());
  return RB_OK;
}

// the deal of a peer-connected plan's cells (cell_of() stride code: +world cyclic, -world
// 32-cell blocks), for rb_plan_pack_cells / rb_plan_unpack_cells; 1 when not connected
extern "C" int rb_cell_of(int j, int start, int stride) { return cell_of(j, start, stride); }

extern "C" int rb_plan_peer_deal(const rb_plan *p) {
  return (p && p->dpeer) ? p->Q.deal : 1;
}

extern "C" long long rb_plan_evals_per_sweep(const rb_plan *p) {
  if (!p) return 0;
  const int g = p->d.geometry;
  const bool two = (g == RB_GEOM_SLAB_PLANE_2 || g == RB_GEOM_SLAB_BGND);
  const long long ncell = two ? p->d.Nl / 2 : p->d.Nl;
  return ncell * (long long)p->Ndir * p->d.Nnu;  // per problem
}

static int plan_check_errors(rb_plan *p) {
  const size_t B = p->d.n_problems;
  CK(cudaMemcpyAsync(p->hflags, p->dflags, 3 * B * sizeof(int), cudaMemcpyDeviceToHost,
                     p->stream));
  CK(cudaStreamSynchronize(p->stream));
  int rc = 0;
  for (size_t b = 0; b < B; ++b) rc = std::max(rc, p->hflags[2 * B + b]);
  return rc;
}

static int direct_sweep(rb_plan *p, bool sharded) {
  if (sharded) return rb_plan_sweep_sharded(p);
  int rc = rb_plan_sweep_local(p, 0, 1);
  if (!rc) rc = rb_plan_finish_sweep_async(p);
  return rc;
}

static int solve_loop(rb_plan *p, rb_info *info, bool sharded) {
  if (!p) return RB_ERR_BAD_ARG;
  PLAN_DEV(p);
  const auto t0 = std::chrono::steady_clock::now();
  const size_t B = p->d.n_problems;
  cudaEventRecord(p->ev[4], p->stream);
  int rc = rb_plan_thin(p);
  if (rc) return rc;
  rc = plan_check_errors(p);
  if (rc) return rc;
  // Sweeps are enqueued up to kDepth ahead of the convergence read-back: the
  // device-side `active` flags turn the kernels of any sweep enqueued after a
  // problem converged into immediate returns, so the iterates are exactly those
  // of a sweep-by-sweep loop while the host never stalls the stream.
  const int kDepth = 4;
  const char *genv = getenv("RB_NO_GRAPHS");  // testing knob: direct launches only
  const bool use_graphs = !(genv && atoi(genv) != 0);
  long long evals = 0;
  int n_active = (int)B, active_during = (int)B;
  bool stop = false;
  for (;;) {
    if (!stop && p->sweeps_issued - p->sweeps_polled < kDepth) {
      if (use_graphs && p->graphs_ok && p->sweeps_issued > 0 &&
          live_problems(p) == p->direct_live) {
        // (the first sweep, and the first one after the launch shapes changed with the number
        // of live problems, goes out directly: it builds the cached work lists -- a host
        // copy that cannot be captured --, sets the kernels' shared-memory attributes and
        // gives the stage times of one sweep)
        rc = sweep_graph(p, sharded);
        if (rc && !p->graphs_ok) {  // capture not possible: direct launches from here on
          rc = direct_sweep(p, sharded);
        }
      } else {
        const int live_now = live_problems(p);
        rc = direct_sweep(p, sharded);
        if (!rc) p->direct_live = live_now;
      }
      if (rc) return rc;
      continue;
    }
    if (p->sweeps_polled >= p->sweeps_issued) break;
    rc = rb_plan_poll_sweep(p, &n_active);
    if (rc) return rc;
    evals += rb_plan_evals_per_sweep(p) * active_during;
    active_during = n_active;
    p->active_hint = n_active;
    if (n_active == 0) stop = true;
  }
  cudaEventRecord(p->ev[5], p->stream);
  rc = plan_check_errors(p);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, p->ev[4], p->ev[5]);
  p->info.ms_device = ms;
  p->info.evals = evals;
  int itmax = 0;
  for (size_t b = 0; b < B; ++b) itmax = std::max(itmax, p->hflags[B + b]);
  p->info.iters = itmax;
  {
    std::vector<unsigned long long> r(B);
    cudaMemcpy(r.data(), p->dresid, B * 8, cudaMemcpyDeviceToHost);
    double mx = 0.0;
    for (size_t b = 0; b < B; ++b) {
      double v;
      memcpy(&v, &r[b], 8);
      mx = std::max(mx, v);
    }
    p->info.resid = mx;
  }
  p->info.ms_total =
      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (info) *info = p->info;
  return rc;
}

extern "C" int rb_plan_solve(rb_plan *p, rb_info *info) { return solve_loop(p, info, false); }
extern "C" int rb_plan_solve_sharded(rb_plan *p, rb_info *info) {
  if (!p || p->Q.world < 2) return RB_ERR_BAD_ARG;
  return solve_loop(p, info, true);
}

// `graphs` != 0: after the first step (direct launches: work lists, stage times) every step
// is one CUDA-graph launch, as in rb_plan_solve -- what a short sharded sweep needs, whose
// seven launches cost as much as its kernels.  Stage times are then those of direct steps.
static int bench_steps(rb_plan *p, int steps, size_t flush_bytes, double *ms_sum, int graphs) {
  if (!p || steps < 0 || !ms_sum) return RB_ERR_BAD_ARG;
  PLAN_DEV(p);
  void *flush = nullptr;
  if (flush_bytes > 0 && cudaMalloc(&flush, flush_bytes) != cudaSuccess) {
    cudaGetLastError();
    return RB_ERR_NOMEM;
  }
  // Every step is bracketed by its own pair of events in the stream (the L2 flush lies
  // between the pairs); the steps are enqueued up to four ahead of the convergence
  // read-backs like in solve_loop, so that the host's latency -- at N > 1 the slowest
  // rank's, handed on by the cross-GPU barrier -- is not part of what is timed.
  const int kDepth = 4;
  const bool sharded = p->Q.world > 1;
  std::vector<cudaEvent_t> ev((size_t)2 * steps, nullptr);
  int rc = RB_OK;
  for (auto &e : ev)
    if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); rc = RB_ERR_CUDA; break; }
  int issued = 0, polled = 0, n_active = 0;
  while (rc == RB_OK && polled < steps) {
    if (issued < steps && p->sweeps_issued - p->sweeps_polled < kDepth) {
      if (flush) cudaMemsetAsync(flush, issued & 0xff, flush_bytes, p->stream);
      cudaEventRecord(ev[2 * issued], p->stream);
      if (graphs && p->graphs_ok && p->sweeps_issued > 0 && p->d.n_problems == 1) {
        rc = sweep_graph(p, sharded);
        if (rc && !p->graphs_ok) rc = direct_sweep(p, sharded);
      } else {
        rc = direct_sweep(p, sharded);
      }
      cudaEventRecord(ev[2 * issued + 1], p->stream);
      ++issued;
      continue;
    }
    rc = rb_plan_poll_sweep(p, &n_active);
    ++polled;
  }
  if (cudaStreamSynchronize(p->stream) != cudaSuccess && rc == RB_OK) rc = RB_ERR_CUDA;
  double sum = 0.0;
  for (int s = 0; s < issued && rc == RB_OK; ++s) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev[2 * s], ev[2 * s + 1]) != cudaSuccess) {
      cudaGetLastError();
      rc = RB_ERR_CUDA;
    }
    sum += ms;
  }
  for (auto &e : ev)
    if (e) cudaEventDestroy(e);
  if (flush) cudaFree(flush);
  *ms_sum = sum;
  return rc;
}

extern "C" int rbi_plan_bench_steps(rb_plan *p, int steps, size_t flush_bytes, int graphs,
                                    double *ms_sum) {
  return bench_steps(p, steps, flush_bytes, ms_sum, graphs);
}

extern "C" int rb_plan_last_info(const rb_plan *p, rb_info *info) {
  if (!p || !info) return RB_ERR_BAD_ARG;
  *info = p->info;
  return RB_OK;
}

extern "C" int rb_plan_get_problem(rb_plan *p, int b, double *T, double *xH1, double *xH2,
                                   double *xHe1, double *xHe2, double *xHe3, double *H1i,
                                   double *He1i, double *He2i, double *H1h, double *He1h,
                                   double *He2h, int *iters) {
  if (!p || b < 0 || b >= p->d.n_problems) return RB_ERR_BAD_ARG;
  PLAN_DEV(p);
  const size_t B = p->d.n_problems, Nl = p->d.Nl, NC = B * Nl;
  double *h = p->hout;  // 12 fields of Nl
  cudaStream_t s = p->stream;
  if (B == 1) {  // [5][Nl] and [6][Nl] are contiguous: three copies instead of twelve
    CK(cudaMemcpyAsync(h, p->dx, 5 * Nl * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h + 5 * Nl, p->dT, Nl * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h + 6 * Nl, p->dsrc, 6 * Nl * 8, cudaMemcpyDeviceToHost, s));
  } else {
    for (int f = 0; f < 5; ++f)
      CK(cudaMemcpyAsync(h + f * Nl, p->dx + f * NC + b * Nl, Nl * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h + 5 * Nl, p->dT + b * Nl, Nl * 8, cudaMemcpyDeviceToHost, s));
    for (int f = 0; f < 6; ++f)
      CK(cudaMemcpyAsync(h + (6 + f) * Nl, p->dsrc + f * NC + b * Nl, Nl * 8,
                         cudaMemcpyDeviceToHost, s));
  }
  CK(cudaMemcpyAsync(p->hflags + B + b, p->dflags + B + b, sizeof(int),
                     cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  double *outs[12] = {xH1, xH2, xHe1, xHe2, xHe3, T, H1i, He1i, He2i, H1h, He1h, He2h};
  const bool rev = p->reversed[b];
  for (int f = 0; f < 12; ++f) {
    if (!outs[f]) continue;
    for (size_t i = 0; i < Nl; ++i) outs[f][i] = h[f * Nl + (rev ? Nl - 1 - i : i)];
  }
  if (iters) *iters = p->hflags[B + b];
  return RB_OK;
}

extern "C" int rb_plan_get_rec(rb_plan *p, int b, double *H1i, double *He1i, double *He2i,
                               double *H1h, double *He1h, double *He2h) {
  if (!p || b < 0 || b >= p->d.n_problems) return RB_ERR_BAD_ARG;
  PLAN_DEV(p);
  const size_t B = p->d.n_problems, Nl = p->d.Nl, NC = B * Nl;
  double *h = p->hout;  // 12*Nl doubles of pinned staging; 6*Nl used
  if (B == 1) {
    CK(cudaMemcpyAsync(h, p->drec, 6 * Nl * 8, cudaMemcpyDeviceToHost, p->stream));
  } else {
    for (int f = 0; f < 6; ++f)
      CK(cudaMemcpyAsync(h + f * Nl, p->drec + f * NC + b * Nl, Nl * 8, cudaMemcpyDeviceToHost,
                         p->stream));
  }
  CK(cudaStreamSynchronize(p->stream));
  double *outs[6] = {H1i, He1i, He2i, H1h, He1h, He2h};
  const bool rev = p->reversed[b];
  for (int f = 0; f < 6; ++f) {
    if (!outs[f]) continue;
    for (size_t i = 0; i < Nl; ++i) outs[f][i] = h[f * Nl + (rev ? Nl - 1 - i : i)];
  }
  return RB_OK;
}

// A large, freshly allocated host buffer is about to be written for the first time (the
// results of a parameter sweep: 403 MB at C5): ask for huge pages, so that the first touch
// costs one fault per 2 MB instead of one per 4 KB (0.1 s, with outliers of 0.5-0.8 s, of page
// faults inside the D2H copy otherwise).  Advisory: ignored where transparent huge pages are off.
static void advise_huge(void *ptr, size_t bytes) {
#ifdef MADV_HUGEPAGE
  if (bytes < ((size_t)16 << 20)) return;
  const size_t pg = (size_t)2 << 20;
  const uintptr_t a = ((uintptr_t)ptr + pg - 1) / pg * pg, e = ((uintptr_t)ptr + bytes) / pg * pg;
  if (e > a) madvise((void *)a, e - a, MADV_HUGEPAGE);
#else
  (void)ptr; (void)bytes;
#endif
}

extern "C" int rb_plan_get_all(rb_plan *p, double *out12, double *rec6, int *iters) {
  if (!p || !out12) return RB_ERR_BAD_ARG;
  PLAN_DEV(p);
  const size_t B = p->d.n_problems, Nl = p->d.Nl, NC = B * Nl;
  const int nf = rec6 ? 18 : 12;
  advise_huge(out12, 12 * NC * 8);
  if (rec6) advise_huge(rec6, 6 * NC * 8);
  DevBuf tmp;
  int rc = tmp.alloc((size_t)nf * NC);
  if (rc) return rc;
  dim3 g((unsigned)((Nl + 255) / 256), (unsigned)B, (unsigned)nf);
  k_gather_outputs<<<g, 256, 0, p->stream>>>(p->P, p->dreversed, nf, tmp.d);
  p->info.n_launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out12, tmp.d, 12 * NC * 8, cudaMemcpyDeviceToHost, p->stream));
  if (rec6)
    CK(cudaMemcpyAsync(rec6, tmp.d + 12 * NC, 6 * NC * 8, cudaMemcpyDeviceToHost, p->stream));
  if (iters)
    CK(cudaMemcpyAsync(iters, p->dflags + B, B * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  return RB_OK;
}

// Results of the whole batch in a PINNED host buffer owned by the plan: the D2H copy runs at
// PCIe speed into resident pages (a copy into freshly allocated pageable memory is dominated
// by first-touch page faults: 0.1 s, with outliers up to 0.8 s, for the 403 MB of C5).
// rb_plan_reserve_results allocates the buffer (call it with the plan: ~0.1 s for 400 MB);
// rb_plan_get_all_pinned gathers + copies and hands out pointers valid until the next call
// or rb_plan_destroy: out12 [12][B][Nl], rec6 [6][B][Nl], iters [B] (any may be NULL).
extern "C" int rb_plan_reserve_results(rb_plan *p) {
  if (!p) return RB_ERR_BAD_ARG;
  PLAN_DEV(p);
  const size_t B = p->d.n_problems, NC = B * p->d.Nl;
  const size_t need = 18 * NC * 8 + ((B * sizeof(int) + 7) / 8) * 8;
  if (p->hall && p->dall && p->hall_bytes >= need) return RB_OK;
  if (p->hall) cudaFreeHost(p->hall);
  cudaFree(p->dall);
  p->hall = nullptr; p->dall = nullptr; p->hall_bytes = 0;
  if (cudaMallocHost((void **)&p->hall, need) != cudaSuccess ||
      cudaMalloc((void **)&p->dall, 18 * NC * 8) != cudaSuccess) {
    cudaGetLastError();
    if (p->hall) cudaFreeHost(p->hall);
    p->hall = nullptr;
    return RB_ERR_NOMEM;
  }
  p->hall_bytes = need;
  return RB_OK;
}

extern "C" int rb_plan_get_all_pinned(rb_plan *p, const double **out12, const double **rec6,
                                      const int **iters) {
  if (!p) return RB_ERR_BAD_ARG;
  int rc = rb_plan_reserve_results(p);
  if (rc) return rc;
  PLAN_DEV(p);
  const size_t B = p->d.n_problems, Nl = p->d.Nl, NC = B * Nl;
  const unsigned nf = rec6 ? 18u : 12u;   // the recombination-photon rates only on request
  dim3 g((unsigned)((Nl + 255) / 256), (unsigned)B, nf);
  k_gather_outputs<<<g, 256, 0, p->stream>>>(p->P, p->dreversed, (int)nf, p->dall);
  p->info.n_launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(p->hall, p->dall, (size_t)nf * NC * 8, cudaMemcpyDeviceToHost, p->stream));
  int *hit = (int *)(p->hall + 18 * NC);
  CK(cudaMemcpyAsync(hit, p->dflags + B, B * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  if (out12) *out12 = p->hall;
  if (rec6) *rec6 = p->hall + 12 * NC;
  if (iters) *iters = hit;
  return RB_OK;
}

extern "C" void *rb_plan_field_ptr(rb_plan *p, int field) {
  if (!p || field < 0 || field > 11) return nullptr;
  const size_t NC = cells(p);
  if (field < 5) return p->dx + field * NC;
  if (field == 5) return p->dT;
  return p->dsrc + (field - 6) * NC;
}
extern "C" void *rb_plan_stream(rb_plan *p) { return p ? (void *)p->stream : nullptr; }
extern "C" int rb_plan_sync(rb_plan *p) {
  if (!p) return RB_ERR_BAD_ARG;
  PLAN_DEV(p);
  CK(cudaStreamSynchronize(p->stream));
  return RB_OK;
}

// ---------------------------------------------------------------------------
// host-pointer entry points (the f2py boundary)
// ---------------------------------------------------------------------------
// i_rec_meth: spheres {1 fixed, 2 outward, 3 radial, 4 isotropic} (sphere_base.py:84-91),
// slabs {1 fixed, 2 ray} (slab_base.py:97-100); only Verner 96 / Hui-Gnedin 97 fits exist
static int check_fits(int geom, int i_rec_meth, int i_photo_fit, int i_rate_fit) {
  const bool sphere = geom == RB_GEOM_SPHERE_BGND || geom == RB_GEOM_SPHERE_STROMGREN;
  if (i_rec_meth < 1 || i_rec_meth > (sphere ? 4 : 2)) return RB_ERR_BAD_ARG;
  if (i_photo_fit != 1 || i_rate_fit != 1) return RB_ERR_UNSUPPORTED;
  return RB_OK;
}

// ---------------------------------------------------------------------------
// plan cache for the one-shot entry points: a solver class calls rb_*_solve once
// per solve, typically many times with the same grid shape; creating a plan
// (15 cudaMalloc, 9 cudaMallocHost, stream, 48 events) costs more than a small
// solve, so released plans are kept and re-armed when the shape matches.
// ---------------------------------------------------------------------------
#include <mutex>
static std::mutex g_cache_mu;
static std::vector<rb_plan *> g_cache;
static const size_t kCacheMax = 8;

static int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) cudaGetLastError();
  return dev;
}

static int plan_acquire(rb_plan **out, const rb_plan_desc &d) {
  const int dev = d.device >= 0 ? d.device : current_device();
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (size_t i = 0; i < g_cache.size(); ++i) {
      rb_plan *p = g_cache[i];
      if (p->d.geometry == d.geometry && p->d.n_problems == d.n_problems && p->d.Nl == d.Nl &&
          p->d.Nnu == d.Nnu && p->d.Nmu == d.Nmu && p->d.device == dev &&
          p->rec_meth == (d.i_rec_meth == 0 ? 1 : d.i_rec_meth)) {
        g_cache.erase(g_cache.begin() + i);
        p->d.i_find_Teq = d.i_find_Teq; p->d.fixed_fcA = d.fixed_fcA; p->d.tol = d.tol;
        p->d.max_sweeps = d.max_sweeps;
        p->P.find_Teq = d.i_find_Teq; p->P.max_sweeps = d.max_sweeps;
        p->P.libm_pow = 0;
        p->P.fcA = (p->rec_meth == 1) ? d.fixed_fcA : 1.0;
        p->P.tol = d.tol; p->P.tol_inner = d.tol * RB_TOL_FRAC;
        for (int X = 0; X < 3; ++X) p->K.em_fac[X] = d.em_fac_set ? d.em_fac[X] : 1.0;
        // (captured graphs are keyed on everything re-armed here: sweep_graph's GraphKey)
        *out = p;
        return RB_OK;
      }
    }
  }
  rb_plan_desc dd = d;
  dd.device = dev;
  return rb_plan_create(out, &dd);
}

static void plan_release(rb_plan *p) {
  if (!p) return;
  rb_plan *evict = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_cache.push_back(p);
    if (g_cache.size() > kCacheMax) {
      evict = g_cache.front();
      g_cache.erase(g_cache.begin());
    }
  }
  if (evict) rb_plan_destroy(evict);
}

extern "C" void rb_cache_clear(void) {
  std::vector<rb_plan *> all;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    all.swap(g_cache);
  }
  for (rb_plan *p : all) rb_plan_destroy(p);
  std::lock_guard<std::mutex> lk(g_teq_mu);
  g_teq.clear();   // tables still bound to live plans stay alive through their shared_ptr
}

static int solve_one(int geom, const double *Edges, const double *nH, const double *nHe,
                     double *T, int Nmu, const double *E_eV, const double *shape,
                     double fixed_fcA, int i_find_Teq, int i_thin, double z, double Hz,
                     double tol, int Nl, int Nnu, double *outs[17], rb_info *info,
                     int i_rec_meth = 1, const double *em_fac = nullptr) {
  const auto t0 = std::chrono::steady_clock::now();
  rb_plan_desc d;
  memset(&d, 0, sizeof(d));
  d.geometry = geom; d.n_problems = 1; d.Nl = Nl; d.Nnu = Nnu; d.Nmu = Nmu;
  d.i_find_Teq = i_find_Teq; d.fixed_fcA = fixed_fcA; d.tol = tol; d.device = -1;
  d.i_rec_meth = i_rec_meth;
  for (int X = 0; X < 3; ++X) d.em_fac[X] = em_fac ? em_fac[X] : 1.0;
  d.em_fac_set = 1;
  rb_plan *p = nullptr;
  int rc = plan_acquire(&p, d);
  if (rc) return rc;
  PLAN_DEV(p);
  {
    // sweep order of the one-shot SphereBgnd entry: RB_SPHERE_BGND_ORDER=gs selects the
    // reference's sequential order (rb_plan_set_sweep_order); cached plans are re-armed
    const char *o = getenv("RB_SPHERE_BGND_ORDER");
    const bool gs = geom == RB_GEOM_SPHERE_BGND && o && (o[0] == 'g' || o[0] == 'G');
    p->sweep_order = gs ? RB_ORDER_GAUSS_SEIDEL : RB_ORDER_JACOBI;
  }
  rc = rb_plan_set_problem(p, 0, Edges, nH, nHe, T, E_eV, shape, z, Hz);
  if (!rc) rc = rb_plan_upload(p);
  if (!rc) {
    if (i_thin == 1) {
      rc = rb_plan_thin(p);
      if (!rc) rc = plan_check_errors(p);
    } else {
      rc = rb_plan_solve(p, nullptr);
    }
  }
  if (!rc) {
    // T is intent(inout): with find_Teq it returns the equilibrium temperature;
    // the two-sided slabs mirror T unconditionally.
    rc = rb_plan_get_problem(p, 0, T, outs[0], outs[1], outs[2], outs[3], outs[4], outs[5],
                             outs[6], outs[7], outs[11], outs[12], outs[13], nullptr);
    if (!rc)  // *_rec outputs (zeros for rec_meth = fixed)
      rc = rb_plan_get_rec(p, 0, outs[8], outs[9], outs[10], outs[14], outs[15], outs[16]);
  }
  if (info) {
    *info = p->info;
    info->ms_total =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  if (rc == RB_OK) plan_release(p);
  else rb_plan_destroy(p);
  return rc;
}

#define PACK17                                                                          \
  double *outs[17] = {xH1,     xH2,      xHe1,     xHe2,    xHe3,     H1i_src,          \
                      He1i_src, He2i_src, H1i_rec, He1i_rec, He2i_rec, H1h_src,         \
                      He1h_src, He2h_src, H1h_rec, He1h_rec, He2h_rec}

extern "C" int rb_sphere_bgnd_solve(
    double *Edges, double *nH, double *nHe, double *T, int Nmu, const double *E_eV,
    const double *shape, int i_rec_meth, double fixed_fcA, int i_photo_fit, int i_rate_fit,
    int i_find_Teq, int i_thin, double em_H1_fac, double em_He1_fac, double em_He2_fac,
    double z, double Hz, double tol, int Nl, int Nnu, double *xH1, double *xH2, double *xHe1,
    double *xHe2, double *xHe3, double *H1i_src, double *He1i_src, double *He2i_src,
    double *H1i_rec, double *He1i_rec, double *He2i_rec, double *H1h_src, double *He1h_src,
    double *He2h_src, double *H1h_rec, double *He1h_rec, double *He2h_rec, rb_info *info) {
  const double em_fac[3] = {em_H1_fac, em_He1_fac, em_He2_fac};
  int rc = check_fits(RB_GEOM_SPHERE_BGND, i_rec_meth, i_photo_fit, i_rate_fit);
  if (rc) return rc;
  PACK17;
  return solve_one(RB_GEOM_SPHERE_BGND, Edges, nH, nHe, T, Nmu, E_eV, shape, fixed_fcA,
                   i_find_Teq, i_thin, z, Hz, tol, Nl, Nnu, outs, info, i_rec_meth, em_fac);
}

extern "C" int rb_sphere_stromgren_solve(
    double *Edges, double *nH, double *nHe, double *T, int Nmu, const double *E_eV,
    const double *shape, int i_rec_meth, double fixed_fcA, int i_photo_fit, int i_rate_fit,
    int i_find_Teq, int i_thin, double em_H1_fac, double em_He1_fac, double em_He2_fac,
    double z, double Hz, double tol, int Nl, int Nnu, double *xH1, double *xH2, double *xHe1,
    double *xHe2, double *xHe3, double *H1i_src, double *He1i_src, double *He2i_src,
    double *H1i_rec, double *He1i_rec, double *He2i_rec, double *H1h_src, double *He1h_src,
    double *He2h_src, double *H1h_rec, double *He1h_rec, double *He2h_rec, rb_info *info) {
  const double em_fac[3] = {em_H1_fac, em_He1_fac, em_He2_fac};
  int rc = check_fits(RB_GEOM_SPHERE_STROMGREN, i_rec_meth, i_photo_fit, i_rate_fit);
  if (rc) return rc;
  PACK17;
  return solve_one(RB_GEOM_SPHERE_STROMGREN, Edges, nH, nHe, T, Nmu, E_eV, shape, fixed_fcA,
                   i_find_Teq, i_thin, z, Hz, tol, Nl, Nnu, outs, info, i_rec_meth, em_fac);
}

extern "C" int rb_slab_bgnd_solve(
    const double *Edges, const double *nH, const double *nHe, double *T, const double *E_eV,
    const double *shape, int i_rec_meth, double fixed_fcA, int i_photo_fit, int i_rate_fit,
    int i_find_Teq, int i_thin, double z, double Hz, double tol, int Nl, int Nnu, double *xH1,
    double *xH2, double *xHe1, double *xHe2, double *xHe3, double *H1i_src, double *He1i_src,
    double *He2i_src, double *H1i_rec, double *He1i_rec, double *He2i_rec, double *H1h_src,
    double *He1h_src, double *He2h_src, double *H1h_rec, double *He1h_rec, double *He2h_rec,
    rb_info *info) {
  int rc = check_fits(RB_GEOM_SLAB_BGND, i_rec_meth, i_photo_fit, i_rate_fit);
  if (rc) return rc;
  PACK17;
  return solve_one(RB_GEOM_SLAB_BGND, Edges, nH, nHe, T, 0, E_eV, shape, fixed_fcA, i_find_Teq,
                   i_thin, z, Hz, tol, Nl, Nnu, outs, info, i_rec_meth);
}

extern "C" int rb_slab_plane_solve(
    const double *Edges, const double *nH, const double *nHe, double *T, const double *E_eV,
    const double *shape, int i_2side, int i_rec_meth, double fixed_fcA, int i_photo_fit,
    int i_rate_fit, int i_find_Teq, int i_thin, double z, double Hz, double tol, int Nl,
    int Nnu, double *xH1, double *xH2, double *xHe1, double *xHe2, double *xHe3,
    double *H1i_src, double *He1i_src, double *He2i_src, double *H1i_rec, double *He1i_rec,
    double *He2i_rec, double *H1h_src, double *He1h_src, double *He2h_src, double *H1h_rec,
    double *He1h_rec, double *He2h_rec, rb_info *info) {
  int rc = check_fits(RB_GEOM_SLAB_PLANE_1, i_rec_meth, i_photo_fit, i_rate_fit);
  if (rc) return rc;
  PACK17;
  return solve_one(i_2side == 1 ? RB_GEOM_SLAB_PLANE_2 : RB_GEOM_SLAB_PLANE_1, Edges, nH, nHe,
                   T, 0, E_eV, shape, fixed_fcA, i_find_Teq, i_thin, z, Hz, tol, Nl, Nnu, outs,
                   info, i_rec_meth);
}

// ---------------------------------------------------------------------------
// small device-buffer helper for the remaining entry points
// ---------------------------------------------------------------------------

// device time of the last single-zone batch kernel of this thread (rbi_last_zone_kernel_ms)
static thread_local double g_zone_kernel_ms = 0.0;
struct ZoneTimer {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  ZoneTimer() {
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
  }
  void stop() {
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) g_zone_kernel_ms = ms;
    cudaGetLastError();
  }
  ~ZoneTimer() { cudaEventDestroy(e0); cudaEventDestroy(e1); }
};
extern "C" double rbi_last_zone_kernel_ms(void) { return g_zone_kernel_ms; }

static int need_device() {
  if (rb_device_count() < 1) {
    snprintf(g_cuda_msg, sizeof(g_cuda_msg), "no CUDA device visible (there is no CPU fallback)");
    return RB_ERR_CUDA;
  }
  return RB_OK;
}

extern "C" int rb_set_column_vs_impact(const double *xH1, const double *xH2, const double *xHe1,
                                       const double *xHe2, const double *xHe3,
                                       const double *Edges, const double *nH, const double *nHe,
                                       const double *T, int Nl, double *NH, double *NHe,
                                       double *NH1, double *NHe1, double *NHe2) {
  (void)xH2; (void)xHe3; (void)T;
  if (Nl < 1) return RB_ERR_BAD_ARG;
  int rc = need_device();
  if (rc) return rc;
  bool rev = false;
  if (mono_inc(Edges, Nl + 1)) rev = true;          // sphere_base.f90:1150-1160
  else if (!mono_dec(Edges, Nl + 1)) return RB_ERR_EDGES_NOT_MONOTONIC;
  const size_t n = Nl;
  std::vector<double> h(6 * n + 1);
  for (int i = 0; i <= Nl; ++i) h[i] = Edges[rev ? Nl - i : i];
  const double *src[5] = {nH, nHe, xH1, xHe1, xHe2};
  for (int f = 0; f < 5; ++f)
    for (int i = 0; i < Nl; ++i) h[(n + 1) + f * n + i] = src[f][rev ? Nl - 1 - i : i];
  DevBuf in, out;
  if ((rc = in.alloc(6 * n + 1))) return rc;
  if ((rc = out.alloc(5 * n + 1))) return rc;
  CK(cudaMemcpy(in.d, h.data(), (6 * n + 1) * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(out.d + 5 * n, 0, 8));
  const double *dE = in.d, *dnH = in.d + n + 1, *dnHe = dnH + n, *dx1 = dnHe + n,
               *dx2 = dx1 + n, *dx3 = dx2 + n;
  k_column_vs_impact<<<(Nl + 127) / 128, 128>>>(dE, dnH, dnHe, dx1, dx2, dx3, Nl, out.d,
                                                out.d + n, out.d + 2 * n, out.d + 3 * n,
                                                out.d + 4 * n, (int *)(out.d + 5 * n));
  CK(cudaGetLastError());
  std::vector<double> o(5 * n + 1);
  CK(cudaMemcpy(o.data(), out.d, (5 * n + 1) * 8, cudaMemcpyDeviceToHost));
  int err;
  memcpy(&err, &o[5 * n], sizeof(int));
  if (err) return err;
  double *dst[5] = {NH, NHe, NH1, NHe1, NHe2};
  for (int f = 0; f < 5; ++f)
    for (int i = 0; i < Nl; ++i) dst[f][i] = o[f * n + (rev ? Nl - 1 - i : i)];
  return RB_OK;
}

// gather `nf` host arrays of nn doubles into one device buffer [nf][nn]
static int upload_fields(DevBuf &buf, const double *const *fields, int nf, int nn) {
  int rc = buf.alloc((size_t)nf * nn);
  if (rc) return rc;
  for (int f = 0; f < nf; ++f)
    CK(cudaMemcpy(buf.d + (size_t)f * nn, fields[f], (size_t)nn * 8, cudaMemcpyHostToDevice));
  return RB_OK;
}

extern "C" int rb_get_kchem(const double *T, const double *fcA_H2, const double *fcA_He2,
                            const double *fcA_He3, int irate, int nn, double *reH2,
                            double *reHe2, double *reHe3, double *ciH1, double *ciHe1,
                            double *ciHe2) {
  if (irate != 1) return RB_ERR_UNSUPPORTED;
  if (nn < 1) return RB_ERR_BAD_ARG;
  int rc = need_device();
  if (rc) return rc;
  DevBuf in, out;
  const double *f[4] = {T, fcA_H2, fcA_He2, fcA_He3};
  if ((rc = upload_fields(in, f, 4, nn))) return rc;
  if ((rc = out.alloc((size_t)6 * nn))) return rc;
  const size_t n = nn;
  k_kchem<<<(nn + 127) / 128, 128>>>(in.d, in.d + n, in.d + 2 * n, in.d + 3 * n, nn, out.d,
                                     out.d + n, out.d + 2 * n, out.d + 3 * n, out.d + 4 * n,
                                     out.d + 5 * n);
  CK(cudaGetLastError());
  double *o[6] = {reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2};
  for (int q = 0; q < 6; ++q) CK(cudaMemcpy(o[q], out.d + q * n, n * 8, cudaMemcpyDeviceToHost));
  return RB_OK;
}

extern "C" int rb_get_kcool(const double *T, const double *fcA_H2, const double *fcA_He2,
                            const double *fcA_He3, int irate, int nn, double *recH2,
                            double *recHe2, double *recHe3, double *cicH1, double *cicHe1,
                            double *cicHe2, double *cecH1, double *cecHe2, double *bremss,
                            double *compton) {
  if (irate != 1) return RB_ERR_UNSUPPORTED;
  if (nn < 1) return RB_ERR_BAD_ARG;
  int rc = need_device();
  if (rc) return rc;
  DevBuf in, out;
  const double *f[4] = {T, fcA_H2, fcA_He2, fcA_He3};
  if ((rc = upload_fields(in, f, 4, nn))) return rc;
  if ((rc = out.alloc((size_t)10 * nn))) return rc;
  const size_t n = nn;
  k_kcool<<<(nn + 127) / 128, 128>>>(in.d, in.d + n, in.d + 2 * n, in.d + 3 * n, nn, out.d);
  CK(cudaGetLastError());
  double *o[10] = {recH2, recHe2, recHe3, cicH1, cicHe1, cicHe2, cecH1, cecHe2, bremss, compton};
  for (int q = 0; q < 10; ++q) CK(cudaMemcpy(o[q], out.d + q * n, n * 8, cudaMemcpyDeviceToHost));
  return RB_OK;
}

// ---------------------------------------------------------------------------
// diagnostics for the unit tests of rb_fastpow.cuh / hg97_eval: the SAME inline source
// evaluated by the host compiler (on_device = 0; runs without a GPU, compared with
// mpmath / the oracle in tests/test_host_logic.py) or by one CUDA thread per element.
// Not reachable from any solver entry point.
// ---------------------------------------------------------------------------
__global__ void k_diag_pow(const double *x, const double *y, int n, double *out) {
  pow_tables_to_smem();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = pow_fast(x[i], y[i]);
}
__global__ void k_diag_hg97(const double *T, const double *a, const double *b, const double *c,
                            int n, double *out, int libm) {
  pow_tables_to_smem();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rb_kchem_t kc;
  rb_kcool_t kl;
  get_kchem_kcool(T[i], a[i], b[i], c[i], kc, kl, libm);
  const double v[16] = {kc.reH2, kc.reHe2, kc.reHe3, kc.ciH1, kc.ciHe1, kc.ciHe2,
                        kl.recH2, kl.recHe2, kl.recHe3, kl.cicH1, kl.cicHe1, kl.cicHe2,
                        kl.cecH1, kl.cecHe2, kl.bremss, kl.compton};
  for (int q = 0; q < 16; ++q) out[(size_t)q * n + i] = v[q];
}

// The device's rate-coefficient evaluation built by the host compiler: bit-identical to what
// the kernels compute (IEEE operations and table reads only, -ffp-contract=off).  Signature
// of the oracle's coefficient hook (oracle/rb_oracle.h: orc_set_coefficient_hook).
extern "C" void rbi_host_coefficients(double T, double fcA_H2, double fcA_He2, double fcA_He3,
                                      double *out16) {
  rb_kchem_t kc;
  rb_kcool_t kl;
  get_kchem_kcool(T, fcA_H2, fcA_He2, fcA_He3, kc, kl, 0);
  const double v[16] = {kc.reH2, kc.reHe2, kc.reHe3, kc.ciH1, kc.ciHe1, kc.ciHe2,
                        kl.recH2, kl.recHe2, kl.recHe3, kl.cicH1, kl.cicHe1, kl.cicHe2,
                        kl.cecH1, kl.cecHe2, kl.bremss, kl.compton};
  for (int q = 0; q < 16; ++q) out16[q] = v[q];
}

// checked build: number of guard bytes behind the plan's device buffers that no longer hold
// the fill pattern (0 = no kernel wrote past the end of a buffer); -1 in the release build
extern "C" int rbi_plan_check_guards(rb_plan *p, char *where, size_t wherelen) {
  if (!p) return -2;
#ifdef RB_DEBUG_CHECKS
  PLAN_DEV(p);
  cudaStreamSynchronize(p->stream);
  int bad = 0;
  std::vector<unsigned char> h(kGuardBytes);
  for (const auto &g : p->guards) {
    if (cudaMemcpy(h.data(), g.first, kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    int n = 0;
    for (unsigned char v : h) n += (v != 0xA5);
    if (n && where && wherelen && bad == 0) snprintf(where, wherelen, "%s", g.second);
    bad += n;
  }
  return bad;
#else
  (void)where; (void)wherelen;
  return -1;
#endif
}

extern "C" int rbi_teq_table_fetch(double fcA, int row0, int n, double *out16, int *depth) {
  if (!out16 || !depth || row0 < 0 || n < 1) return RB_ERR_BAD_ARG;
  int rc = need_device();
  if (rc) return rc;
  std::shared_ptr<TeqTable> t = teq_table_get(fcA, 0);
  if (!t) return RB_ERR_UNSUPPORTED;
  *depth = t->depth;
  if ((long long)row0 + n > ((long long)1 << t->depth) + 1) return RB_ERR_BAD_ARG;
  CK(cudaMemcpy(out16, t->rows + (size_t)row0 * kTeqStride, (size_t)n * kTeqStride * 8,
                cudaMemcpyDeviceToHost));
  return RB_OK;
}

// decisions of rb_ion.cuh ratio_err_exceeds (reciprocal-based, certified) and of the IEEE
// divisions it stands for, for n triples of quotients: out[i] = fast | exact << 1
__global__ void k_diag_ratio_err(const double *a, const double *b, int n, double tol, int *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a1 = a[3 * i], a2 = a[3 * i + 1], a3 = a[3 * i + 2];
  const double b1 = b[3 * i], b2 = b[3 * i + 1], b3 = b[3 * i + 2];
  const int fast = ratio_err_exceeds(a1, b1, a2, b2, a3, b3, tol) ? 1 : 0;
  const int exact = (fabs(fmax(fmax(a1 / b1, a2 / b2), a3 / b3) - 1.0) > tol) ? 1 : 0;
  const int up = (a1 > b1) ? 1 : 0, up_div = ((a1 / b1) > 1.0) ? 1 : 0;
  const int dn = (a1 < b1) ? 1 : 0, dn_div = ((a1 / b1) < 1.0) ? 1 : 0;
  out[i] = fast | (exact << 1) | (up << 2) | (up_div << 3) | (dn << 4) | (dn_div << 5);
}
extern "C" int rbi_diag_ratio_err(const double *a, const double *b, int n, double tol, int *out) {
  if (!a || !b || !out || n < 1) return RB_ERR_BAD_ARG;
  int rc = need_device();
  if (rc) return rc;
  DevBuf da, db, dout;
  if ((rc = da.alloc((size_t)3 * n)) || (rc = db.alloc((size_t)3 * n)) || (rc = dout.alloc((size_t)n)))
    return rc;
  CK(cudaMemcpy(da.d, a, (size_t)3 * n * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db.d, b, (size_t)3 * n * 8, cudaMemcpyHostToDevice));
  k_diag_ratio_err<<<(n + 255) / 256, 256>>>(da.d, db.d, n, tol, (int *)dout.d);
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, dout.d, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
  return RB_OK;
}

extern "C" int rbi_plan_set_far_field(rb_plan *p, int on) {
  if (!p) return RB_ERR_BAD_ARG;
  if (on && !p->dfar) return RB_ERR_UNSUPPORTED;
  p->use_far = on != 0;   // part of the graph key and of the work-list cache key
  return RB_OK;
}

extern "C" int rbi_plan_set_libm_pow(rb_plan *p, int on) {
  if (!p) return RB_ERR_BAD_ARG;
  p->P.libm_pow = on ? 1 : 0;   // part of the graph key: captured sweeps are re-captured
  return RB_OK;
}

extern "C" int rbi_diag_pow(const double *x, const double *y, int n, double *out, int on_device) {
  if (!x || !y || !out || n < 1) return RB_ERR_BAD_ARG;
  if (!on_device) {
    for (int i = 0; i < n; ++i) out[i] = pow_fast(x[i], y[i]);
    return RB_OK;
  }
  int rc = need_device();
  if (rc) return rc;
  DevBuf in, o;
  const double *f[2] = {x, y};
  if ((rc = upload_fields(in, f, 2, n))) return rc;
  if ((rc = o.alloc((size_t)n))) return rc;
  k_diag_pow<<<(n + 127) / 128, 128>>>(in.d, in.d + n, n, o.d);
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, o.d, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return RB_OK;
}

extern "C" int rbi_diag_hg97(const double *T, const double *fcA_H2, const double *fcA_He2,
                             const double *fcA_He3, int n, double *out16, int on_device,
                             int libm) {
  if (!T || !fcA_H2 || !fcA_He2 || !fcA_He3 || !out16 || n < 1) return RB_ERR_BAD_ARG;
  if (!on_device) {
    for (int i = 0; i < n; ++i) {
      rb_kchem_t kc;
      rb_kcool_t kl;
      get_kchem_kcool(T[i], fcA_H2[i], fcA_He2[i], fcA_He3[i], kc, kl, libm);
      const double v[16] = {kc.reH2, kc.reHe2, kc.reHe3, kc.ciH1, kc.ciHe1, kc.ciHe2,
                            kl.recH2, kl.recHe2, kl.recHe3, kl.cicH1, kl.cicHe1, kl.cicHe2,
                            kl.cecH1, kl.cecHe2, kl.bremss, kl.compton};
      for (int q = 0; q < 16; ++q) out16[(size_t)q * n + i] = v[q];
    }
    return RB_OK;
  }
  int rc = need_device();
  if (rc) return rc;
  DevBuf in, o;
  const double *f[4] = {T, fcA_H2, fcA_He2, fcA_He3};
  if ((rc = upload_fields(in, f, 4, n))) return rc;
  if ((rc = o.alloc((size_t)16 * n))) return rc;
  const size_t nn = n;
  k_diag_hg97<<<(n + 127) / 128, 128>>>(in.d, in.d + nn, in.d + 2 * nn, in.d + 3 * nn, n, o.d,
                                        libm);
  CK(cudaGetLastError());
  CK(cudaMemcpy(out16, o.d, (size_t)16 * n * 8, cudaMemcpyDeviceToHost));
  return RB_OK;
}

static int fetch_err(DevBuf &out, size_t off) {
  int err = 0;
  if (cudaMemcpy(&err, out.d + off, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) {
    cudaGetLastError();
    return RB_ERR_CUDA;
  }
  return err;
}

extern "C" int rb_solve_ce(const double *reH2, const double *reHe2, const double *reHe3,
                           const double *ciH1, const double *ciHe1, const double *ciHe2, int nn,
                           double *xH1, double *xH2, double *xHe1, double *xHe2, double *xHe3) {
  if (nn < 1) return RB_ERR_BAD_ARG;
  int rc = need_device();
  if (rc) return rc;
  DevBuf in, out;
  const double *f[6] = {reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2};
  if ((rc = upload_fields(in, f, 6, nn))) return rc;
  const size_t n = nn;
  if ((rc = out.alloc(5 * n + 1))) return rc;
  CK(cudaMemset(out.d + 5 * n, 0, 8));
  k_zone_pce<<<(nn + 127) / 128, 128>>>(in.d, nn, 0.0, out.d, (int *)(out.d + 5 * n), 1);
  CK(cudaGetLastError());
  double *o[5] = {xH1, xH2, xHe1, xHe2, xHe3};
  for (int q = 0; q < 5; ++q) CK(cudaMemcpy(o[q], out.d + q * n, n * 8, cudaMemcpyDeviceToHost));
  return fetch_err(out, 5 * n);
}

template <typename F>
static int launch_lockstep(int nn, F &&launch) {
  int cpt = 1;
  while (cpt * 512 < nn) cpt *= 4;
  if (cpt > 16) return RB_ERR_UNSUPPORTED;  // lock-step batch limit: 8192 zones
  int threads = (nn + cpt - 1) / cpt;
  threads = ((threads + 31) / 32) * 32;
  launch(cpt, threads);
  return RB_OK;
}

extern "C" int rb_solve_pce(const double *nH, const double *nHe, const double *reH2,
                            const double *reHe2, const double *reHe3, const double *ciH1,
                            const double *ciHe1, const double *ciHe2, const double *H1i,
                            const double *He1i, const double *He2i, double tol, int nn,
                            int lockstep, double *xH1, double *xH2, double *xHe1, double *xHe2,
                            double *xHe3) {
  if (nn < 1) return RB_ERR_BAD_ARG;
  int rc = need_device();
  if (rc) return rc;
  DevBuf in, out;
  const double *f[11] = {nH, nHe, reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2, H1i, He1i, He2i};
  if ((rc = upload_fields(in, f, 11, nn))) return rc;
  const size_t n = nn;
  if ((rc = out.alloc(5 * n + 1))) return rc;
  CK(cudaMemset(out.d + 5 * n, 0, 8));
  int *derr = (int *)(out.d + 5 * n);
  if (lockstep) {
    rc = launch_lockstep(nn, [&](int cpt, int threads) {
      switch (cpt) {
        case 1: k_zone_lockstep<1><<<1, threads>>>(in.d, nn, 0, 0, tol, out.d, derr, 0, TeqTab{nullptr, 0, 0}); break;
        case 4: k_zone_lockstep<4><<<1, threads>>>(in.d, nn, 0, 0, tol, out.d, derr, 0, TeqTab{nullptr, 0, 0}); break;
        default: k_zone_lockstep<16><<<1, threads>>>(in.d, nn, 0, 0, tol, out.d, derr, 0, TeqTab{nullptr, 0, 0});
      }
    });
    if (rc) return rc;
  } else {
    ZoneTimer zt;
    k_zone_pce<<<(nn + 127) / 128, 128>>>(in.d, nn, tol, out.d, derr, 0);
    zt.stop();
  }
  CK(cudaGetLastError());
  double *o[5] = {xH1, xH2, xHe1, xHe2, xHe3};
  for (int q = 0; q < 5; ++q) CK(cudaMemcpy(o[q], out.d + q * n, n * 8, cudaMemcpyDeviceToHost));
  return fetch_err(out, 5 * n);
}

extern "C" int rb_solve_pcte(const double *nH, const double *nHe, const double *H1i,
                             const double *He1i, const double *He2i, const double *H1h,
                             const double *He1h, const double *He2h, double z, double Hz,
                             const double *fcA_H2, const double *fcA_He2, const double *fcA_He3,
                             int irate, double tol, int nn, int lockstep, double *xH1,
                             double *xH2, double *xHe1, double *xHe2, double *xHe3, double *T) {
  if (irate != 1) return RB_ERR_UNSUPPORTED;
  if (nn < 1) return RB_ERR_BAD_ARG;
  int rc = need_device();
  if (rc) return rc;
  bool uniform = true;
  for (int i = 1; i < nn && uniform; ++i)
    uniform = fcA_H2[i] == fcA_H2[0] && fcA_He2[i] == fcA_He2[0] && fcA_He3[i] == fcA_He3[0];
  // the lock-step kernel carries one case-A triple for the batch
  if (lockstep && !uniform) return RB_ERR_UNSUPPORTED;
  // one case-A fraction for all zones and ions (the usual call): the memo of the rate fits
  // over the temperature bisection tree applies (rb_ion.cuh: TeqTab)
  std::shared_ptr<TeqTable> tq;
  if (uniform && fcA_He2[0] == fcA_H2[0] && fcA_He3[0] == fcA_H2[0]) tq = teq_table_get(fcA_H2[0], 0);
  const TeqTab tb = teq_tab_of(tq);
  DevBuf in, out;
  const double *f[11] = {nH, nHe, H1i, He1i, He2i, H1h, He1h, He2h, fcA_H2, fcA_He2, fcA_He3};
  if ((rc = upload_fields(in, f, 11, nn))) return rc;
  const size_t n = nn;
  if ((rc = out.alloc(6 * n + 1))) return rc;
  CK(cudaMemset(out.d + 6 * n, 0, 8));
  int *derr = (int *)(out.d + 6 * n);
  if (lockstep) {
    rc = launch_lockstep(nn, [&](int cpt, int threads) {
      switch (cpt) {
        case 1: k_zone_lockstep<1><<<1, threads>>>(in.d, nn, z, Hz, tol, out.d, derr, 1, tb); break;
        case 4: k_zone_lockstep<4><<<1, threads>>>(in.d, nn, z, Hz, tol, out.d, derr, 1, tb); break;
        default: k_zone_lockstep<16><<<1, threads>>>(in.d, nn, z, Hz, tol, out.d, derr, 1, tb);
      }
    });
    if (rc) return rc;
  } else {
    ZoneTimer zt;
    k_zone_pcte<<<(nn + 127) / 128, 128>>>(in.d, nn, z, Hz, tol, out.d, derr, tb);
    zt.stop();
  }
  CK(cudaGetLastError());
  double *o[6] = {xH1, xH2, xHe1, xHe2, xHe3, T};
  for (int q = 0; q < 6; ++q) CK(cudaMemcpy(o[q], out.d + q * n, n * 8, cudaMemcpyDeviceToHost));
  return fetch_err(out, 6 * n);
}

// ---------------------------------------------------------------------------
// micro-benchmarks
// ---------------------------------------------------------------------------
extern "C" int rbi_microbench(int kind, int device, double *ops_per_s) {
  if (!ops_per_s || kind < 0 || kind > 3) return RB_ERR_BAD_ARG;
  int rc = need_device();
  if (rc) return rc;
  DevGuard dev_guard_(device);
  int dev = 0;
  CK(cudaGetDevice(&dev));
  int n_sm = 0;
  CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  // DFMA: 4 CTAs x 8 warps per SM, 64 FMAs per trip, ~60 ms per repetition at 17e12 FMA/s
  const int blocks = n_sm * (kind == 0 ? 4 : 8), threads = 256;
  const int iters = (kind == 0) ? 100000 : (kind == 3 ? 128 : 512);
  const double per_trip = (kind == 0) ? 64.0 : 8.0;
  DevBuf out;
  if ((rc = out.alloc((size_t)blocks * threads))) return rc;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  const int reps = (kind == 0) ? 3 : 4;
  for (int rep = 0; rep < reps; ++rep) {
    cudaEventRecord(e0);
    switch (kind) {
      case 0: k_microbench<0><<<blocks, threads>>>(out.d, iters, 0.5); break;
      case 1: k_microbench<1><<<blocks, threads>>>(out.d, iters, 0.5); break;
      case 2: k_microbench<2><<<blocks, threads>>>(out.d, iters, 0.5); break;
      default: k_microbench<3><<<blocks, threads>>>(out.d, iters, 0.5);
    }
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads * iters * per_trip;
    if (rep > 0) best = std::max(best, ops / (ms * 1e-3));
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  CK(cudaGetLastError());
  *ops_per_s = best;
  return RB_OK;
}
