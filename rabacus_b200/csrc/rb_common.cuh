// rb_common.cuh -- shared definitions for the rabacus_b200 CUDA library.
//
// The whole library is compiled with -fmad=false: plain `a*b+c` expressions are
// evaluated as the reference's gfortran build evaluates them (separate IEEE
// multiply and add, rabacus setup.py:45-48 has no -ffast-math / -march flags),
// which keeps the bisection branch decisions of the ion solver on the same
// path.  Where fusion is wanted (the hot frequency / ray loops) it is written
// explicitly with fma().
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/rabacus_b200.h"

// rabacus/f2py/physical_constants.f90:7-21
#define RB_PI 3.141592653589793e+00
#define RB_KB_ERG_K 1.380648800000000e-16
#define RB_H_ERG_S 6.626069570000000e-27
#define RB_EV_2_ERG 1.602176530000000e-12
#define RB_ERG_2_EV 6.241509479607719e+11

// rabacus/f2py/ion_solver.f90:13-15
#define RB_ION_MAX_ITER 5000
#define RB_TFLOOR 7.5e3
#define RB_TMAX 1.0e5
// sphere_bgnd.f90:45-46 (same in sphere_stromgren.f90:44-45, slab_plane.f90:37-38)
#define RB_OUTER_MAX_ITER 500
#define RB_TOL_FRAC 1.0e-2
// slab_bgnd.f90:153-155,278-282: silent stop after 200 sweeps
#define RB_SLAB_BGND_MAX_SWEEPS 200

// Checked build (rabacus_b200/build.py --checked -> librabacus_b200_checked.so): device-side
// assertions on every computed index of the hot kernels (shared-memory records, far-field
// rows, table offsets, work items, memo rows, peer staging) and guard words behind every
// device buffer of a plan (rbi_plan_check_guards).  compute-sanitizer is closed on the GPU
// pool this was developed on; scratch/sanitize_small.py runs every kernel family on small
// grids under this build instead (profiles/r02_sanitizer.txt).
#ifdef RB_DEBUG_CHECKS
#include <assert.h>
#define RB_ASSERT(c) assert(c)
#else
#define RB_ASSERT(c) ((void)0)
#endif

#define RB_HD __host__ __device__ __forceinline__

// Local cell index j of a rank -> cell il.  Two deals of the swept cells over ranks:
//   stride > 0  cyclic: il = start + j * stride (start = rank, stride = world)
//   stride < 0  32-cell BLOCKS dealt cyclically over W = -stride ranks (start = rank):
//               il = ((j / 32) W + rank) 32 + j % 32 -- the 32 lanes of a column-kernel warp
//               then hold 32 CONSECUTIVE shells whatever the world size, so a shard's ray
//               walks are as compact as a single GPU's (cyclic cells at W = 8 spread a warp
//               over 256 shells and cost 2.5x the instructions per ray); the ray cost still
//               balances over the ranks because the blocks interleave.  Needs
//               (cells to sweep) % (32 W) == 0.
static __host__ __device__ __forceinline__ int cell_of(int j, int start, int stride) {
  return stride > 0 ? start + j * stride : ((((j >> 5) * (-stride)) + start) << 5) + (j & 31);
}
// distance in cells between neighbouring lanes of a warp
static __host__ __device__ __forceinline__ int lane_stride_of(int stride) {
  return stride > 0 ? stride : 1;
}
// heavy bodies (dozens of pow/exp): real calls keep code size and compile time sane
#define RB_HDN static __host__ __device__ __noinline__

struct rb_kchem_t {
  double reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2;
};
struct rb_kcool_t {
  double recH2, recHe2, recHe3, cicH1, cicHe1, cicHe2, cecH1, cecHe2, bremss,
      compton;
};
struct rb_frac_t {
  double xH1, xH2, xHe1, xHe2, xHe3;
};
