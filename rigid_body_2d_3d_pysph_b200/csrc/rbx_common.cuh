// Shared device helpers for librbx (sm_100a).  See include/rbx.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rbx.h"

#define RBX_CHUNK 128     // threads per CTA = max particles per work item
#define RBX_TILE 768      // staged source particles per shared-memory tile

#define RBX_CHECK_LAUNCH()                                   \
  do {                                                       \
    if (cudaPeekAtLastError() != cudaSuccess) {              \
      cudaGetLastError();                                    \
      return RBX_ERR_LAUNCH;                                 \
    }                                                        \
  } while (0)

static inline int rbx_blocks(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  return (int)(b < 1 ? 1 : b);
}

// SMs of the current device (launch geometry of the persistent kernels);
// a device attribute, cached per device.
static inline int rbx_sm_count() {
  static int cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

// ---- ordered-uint64 encoding of doubles, for atomicMin/atomicMax ----
__device__ __forceinline__ unsigned long long rbx_ord(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double rbx_unord(unsigned long long k) {
  unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// r2 = dx*dx + dy*dy + dz*dz with every product and sum rounded separately
// (no FMA contraction): the neighbour predicate has to be bit-exact against
// the CPU path (SURVEY.md App. C-1, section 7 "Bit-exact neighbour sets").
__device__ __forceinline__ double rbx_r2(double dx, double dy, double dz) {
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
// (k h)^2 = k*k*h*h evaluated left to right
__device__ __forceinline__ double rbx_h2(double rs2, double h) {
  return __dmul_rn(__dmul_rn(rs2, h), h);
}

// QuinticSpline.kernel [upstream pysph.base.kernels; SURVEY App. C-3]
template <int DIM>
__device__ __forceinline__ double rbx_quintic(double rij, double h) {
  const double M_1_PI_ = 0.31830988618379067154;
  double h1 = 1. / h;
  double q = rij * h1;
  double fac;
  if (DIM == 2) fac = (M_1_PI_ * 7.0 / 478.0) * h1 * h1;
  else if (DIM == 3) fac = (M_1_PI_ / 120.0) * h1 * h1 * h1;
  else fac = (1.0 / 120.0) * h1;
  double t3 = 3. - q, t2 = 2. - q, t1 = 1. - q;
  double val;
  if (q > 3.0) val = 0.0;
  else {
    // x^5 as (x^2)^2 x: three multiplies instead of four (the CPU path
    // multiplies left to right; the difference is one rounding, 1e-16)
    const double a3 = t3 * t3, a2 = t2 * t2, a1 = t1 * t1;
    val = a3 * a3 * t3;
    if (q <= 2.0) val -= 6.0 * (a2 * a2 * t2);
    if (q <= 1.0) val += 15. * (a1 * a1 * t1);
  }
  return val * fac;
}

// The same kernel without a branch (clamped terms are exact zeros, every
// other operation is the one above): straight-line code that the scheduler
// can interleave with a second, independent evaluation.
template <int DIM>
__device__ __forceinline__ double rbx_quintic_nb(double rij, double h) {
  const double M_1_PI_ = 0.31830988618379067154;
  const double h1 = 1. / h;
  const double q = rij * h1;
  double fac;
  if (DIM == 2) fac = (M_1_PI_ * 7.0 / 478.0) * h1 * h1;
  else if (DIM == 3) fac = (M_1_PI_ / 120.0) * h1 * h1 * h1;
  else fac = (1.0 / 120.0) * h1;
  const double t3 = q < 3. ? 3. - q : 0., t2 = q < 2. ? 2. - q : 0., t1 = q < 1. ? 1. - q : 0.;
  const double a3 = t3 * t3, a2 = t2 * t2, a1 = t1 * t1;
  double val = a3 * a3 * t3;
  val -= 6.0 * (a2 * a2 * t2);
  val += 15. * (a1 * a1 * t1);
  return val * fac;
}

// 1/sqrt(x) without the special-case branch of the library routine:
// MUFU.RSQ64H seed (2^-22) and one third-order step y += y e (1/2 + 3/8 e),
// e = 1 - x y^2 (-> 1e-16 relative), the library's own refinement.
__device__ __forceinline__ double rbx_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y * y, 1.0);
  return fma(fma(e, 0.375, 0.5), y * e, y);
}

// stage1 / stage3 of the body particles (rigid_body_3d.py:62-95, 192-225):
// v = vcm + omega x (R r0).  One definition for k_pose and for the contact
// law, which forms the velocities of the few particles in contact itself
// (RBX_PARAM_BODY_VEL) -- the same instruction sequence, the same bits.
__device__ __forceinline__ void rbx_point_velocity(const double *Rv, const double *om,
                                                   const double *vc, double x0, double y0,
                                                   double z0, double &u, double &v, double &w) {
  const double dx = (Rv[0] * x0 + Rv[1] * y0 + Rv[2] * z0);
  const double dy = (Rv[3] * x0 + Rv[4] * y0 + Rv[5] * z0);
  const double dz = (Rv[6] * x0 + Rv[7] * y0 + Rv[8] * z0);
  const double du = om[1] * dz - om[2] * dy;
  const double dv = om[2] * dx - om[0] * dz;
  const double dw = om[0] * dy - om[1] * dx;
  u = vc[0] + du;
  v = vc[1] + dv;
  w = vc[2] + dw;
}

// ---- packed FP32 pairs --------------------------------------------------------
// sm_100 has FFMA2 / FMUL2 / FADD2: one instruction, one issue slot, two
// IEEE round-to-nearest FP32 operations on a 64-bit register pair (PTX
// fma/mul/add.rn.f32x2).  Where the same arithmetic runs on two independent
// values an issue-bound kernel needs half the FP32 instructions.  A pair is
// carried as a 64-bit integer; ptxas folds rbx_f2(a, a) into a broadcast
// operand and immediates into the instruction.
typedef unsigned long long rbx_f2_t;
__device__ __forceinline__ rbx_f2_t rbx_f2(float lo, float hi) {
  rbx_f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void rbx_f2_get(rbx_f2_t v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ rbx_f2_t rbx_f2_fma(rbx_f2_t a, rbx_f2_t b, rbx_f2_t c) {
  rbx_f2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ rbx_f2_t rbx_f2_mul(rbx_f2_t a, rbx_f2_t b) {
  rbx_f2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ rbx_f2_t rbx_f2_add(rbx_f2_t a, rbx_f2_t b) {
  rbx_f2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

__device__ __forceinline__ void rbx_prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ int rbx_cell_coord(double x, double x0, double inv, int n) {
  int c = (int)floor((x - x0) * inv);
  return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

__device__ __forceinline__ double rbx_warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double rbx_warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double rbx_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
