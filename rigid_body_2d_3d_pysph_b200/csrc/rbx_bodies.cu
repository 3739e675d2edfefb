// Per-body kernels: the force/torque reduction (k_reduce: one warp per body,
// lanes stride over the body's particles, fixed shuffle tree => deterministic)
// and the rigid steppers (k_bodies, k_rk2: one thread per body for the 3x3
// algebra), plus the per-particle pose kernel and the multi-GPU halo payload.
//
//   SumUpExternalForces.reduce        rigid_body_common.py:128-175
//   GTVFRigidBody3DStep.py_stage1/3   rigid_body_3d.py:41-60, 171-190
//   GTVFRigidBody3DStep.py_stage2     rigid_body_3d.py:97-132
//   normalize_R_orientation           rigid_body_common.py:178-203
//   GTVFRigidBody2DStep               rigid_body_2d.py:40-205
//   RK2RigidBody3DStep                rigid_body_3d.py:406-575
//   stage1/2/3 particle updates       rigid_body_3d.py:62-95, 134-169, 192-225
#include "rbx_common.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace {

__device__ __forceinline__ void matvec3(const double *A, const double *b, double *o) {
#pragma unroll
  for (int r = 0; r < 3; r++) o[r] = A[3 * r] * b[0] + A[3 * r + 1] * b[1] + A[3 * r + 2] * b[2];
}
__device__ __forceinline__ void matmul3(const double *A, const double *B, double *C) {
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++)
      C[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}

// classical Gram-Schmidt on the columns of the row-major 3x3 `o`
__device__ void normalize_R(double *o) {
  double a1[3] = {o[0], o[3], o[6]}, a2[3] = {o[1], o[4], o[7]}, a3[3] = {o[2], o[5], o[8]};
  double na1 = sqrt(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]);
  double b1[3] = {a1[0] / na1, a1[1] / na1, a1[2] / na1};
  double d12 = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
  double b2[3] = {a2[0] - d12 * b1[0], a2[1] - d12 * b1[1], a2[2] - d12 * b1[2]};
  double nb2 = sqrt(b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2]);
#pragma unroll
  for (int k = 0; k < 3; k++) b2[k] /= nb2;
  double d13 = b1[0] * a3[0] + b1[1] * a3[1] + b1[2] * a3[2];
  double d23 = b2[0] * a3[0] + b2[1] * a3[1] + b2[2] * a3[2];
  double b3[3];
#pragma unroll
  for (int k = 0; k < 3; k++) b3[k] = a3[k] - d13 * b1[k] - d23 * b2[k];
  double nb3 = sqrt(b3[0] * b3[0] + b3[1] * b3[1] + b3[2] * b3[2]);
#pragma unroll
  for (int k = 0; k < 3; k++) b3[k] /= nb3;
  o[0] = b1[0]; o[3] = b1[1]; o[6] = b1[2];
  o[1] = b2[0]; o[4] = b2[1]; o[7] = b2[2];
  o[2] = b3[0]; o[5] = b3[1]; o[8] = b3[2];
}

// R <- GS(Rbase + h [omega]x R);  Iinv_g <- R Iinv_b R^T (skipped if planar)
__device__ void rotate_body(double *R, const double *Rbase, const double *om, double h,
                            const double *iinv_b, double *iinv_g) {
  double W[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
  double Rl[9], rd[9];
#pragma unroll
  for (int k = 0; k < 9; k++) Rl[k] = R[k];
  matmul3(W, Rl, rd);
#pragma unroll
  for (int k = 0; k < 9; k++) Rl[k] = Rbase[k] + rd[k] * h;
  normalize_R(Rl);
#pragma unroll
  for (int k = 0; k < 9; k++) R[k] = Rl[k];
  if (iinv_g) {
    double Rt[9] = {Rl[0], Rl[3], Rl[6], Rl[1], Rl[4], Rl[7], Rl[2], Rl[5], Rl[8]};
    double Ib[9], tmp[9], out[9];
#pragma unroll
    for (int k = 0; k < 9; k++) Ib[k] = iinv_b[k];
    matmul3(Rl, Ib, tmp);
    matmul3(tmp, Rt, out);
#pragma unroll
    for (int k = 0; k < 9; k++) iinv_g[k] = out[k];
  }
}

__device__ __forceinline__ void kick_body(const RbxScene &S, int b, double dtb2) {
  const int i3 = 3 * b, i9 = 9 * b;
  const double M = S.total_mass[b];
  if (S.planar) {
#pragma unroll
    for (int j = 0; j < 2; j++) S.vcm[i3 + j] = S.vcm[i3 + j] + (dtb2 * S.force[i3 + j] / M);
    S.omega[i3 + 2] += dtb2 * S.torque[i3 + 2] / S.izz[b];
  } else {
    double L[3], om[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      S.vcm[i3 + j] = S.vcm[i3 + j] + (dtb2 * S.force[i3 + j] / M);
      L[j] = S.ang_mom[i3 + j] + (dtb2 * S.torque[i3 + j]);
      S.ang_mom[i3 + j] = L[j];
    }
    matvec3(S.iinv_g + i9, L, om);
#pragma unroll
    for (int j = 0; j < 3; j++) S.omega[i3 + j] = om[j];
  }
}

// Has body b moved more than skin/2 since the neighbour lists were built?
// |x_p - x_p,ref| = |dxcm + (R - R_ref) r0| <= |dxcm| + |R - R_ref|_F rmax.
__device__ __forceinline__ void check_displacement(const RbxScene &S, int b, double skin) {
  if (!S.rebuild || !S.xcm_ref) return;
  const int i3 = 3 * b, i9 = 9 * b;
  double d2 = 0., r2 = 0.;
#pragma unroll
  for (int j = 0; j < 3; j++) { const double d = S.xcm[i3 + j] - S.xcm_ref[i3 + j]; d2 += d * d; }
#pragma unroll
  for (int k = 0; k < 9; k++) { const double d = S.R[i9 + k] - S.R_ref[i9 + k]; r2 += d * d; }
  const double bound = sqrt(d2) + sqrt(r2) * S.rmax[b];
  if (!(bound <= 0.5 * skin)) atomicOr(S.rebuild, 1u);   // NaN also rebuilds
}

__device__ __forceinline__ void drift_body(const RbxScene &S, int b, double dt) {
  const int i3 = 3 * b, i9 = 9 * b;
  const int nd = S.planar ? 2 : 3;
  for (int j = 0; j < nd; j++) S.xcm[i3 + j] = S.xcm[i3 + j] + dt * S.vcm[i3 + j];
  double om[3] = {S.omega[i3], S.omega[i3 + 1], S.omega[i3 + 2]};
  if (S.R_prev) {
#pragma unroll
    for (int k = 0; k < 9; k++) S.R_prev[i9 + k] = S.R[i9 + k];
  }
  rotate_body(S.R + i9, S.R + i9, om, dt, S.iinv_b + i9, S.planar ? nullptr : S.iinv_g + i9);
}

// SumUpExternalForces.reduce :158-175, a warp per body: lanes stride over the
// body's particles (contiguous), fixed shuffle tree => deterministic.  A
// kernel of its own (few registers, short-lived warps): it streams 48 bytes
// per particle and should run at HBM speed.
__global__ void __launch_bounds__(128, 8) k_reduce(RbxScene S) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= S.n_bodies) return;
  const int b = warp;
  // no particle of this body is in contact: every f_i = m_i g, so
  // F = M g and the torque of a uniform field about the centre of mass is 0
  // (the dense sum gives rounding noise, 1e-16 of its terms)
  if (S.body_tag && S.body_tag[b] == 0) {
    if (lane < 3) {
      S.force[3 * b + lane] = S.total_mass[b] * S.gravity[lane];
      S.torque[3 * b + lane] = 0.;
    }
    return;
  }
  double v6[6] = {0, 0, 0, 0, 0, 0};
  const int q0 = S.chunk_start[S.body_chunk[b]], q1 = S.chunk_start[S.body_chunk[b + 1]];
  const double cx = S.xcm[3 * b], cy = S.xcm[3 * b + 1], cz = S.xcm[3 * b + 2];
#pragma unroll 4
  for (int q = q0 + lane; q < q1; q += 32) {
    const double fx = S.fx[q], fy = S.fy[q], fz = S.fz[q];
    const double dx = S.x[q] - cx, dy = S.y[q] - cy, dz = S.z[q] - cz;
    v6[0] += fx; v6[1] += fy; v6[2] += fz;
    v6[3] += (dy * fz - dz * fy);
    v6[4] += (dz * fx - dx * fz);
    v6[5] += (dx * fy - dy * fx);
  }
#pragma unroll
  for (int a = 0; a < 6; a++) v6[a] = rbx_warp_sum(v6[a]);
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) { S.force[3 * b + a] = v6[a]; S.torque[3 * b + a] = v6[3 + a]; }
  }
}

// mode bits: 2 = kick, 4 = drift, 8 = kick before drift (1 = reduce: k_reduce)
// order executed: [kick (post)] [kick (pre)] [drift]
// The 3x3 algebra of kick and drift (a few hundred dependent FP64
// instructions with divisions and square roots) is one thread's work: a
// thread per body, 32 bodies per warp.
__global__ void k_bodies(RbxScene S, int mode, double dt, double skin) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= S.n_bodies) return;
  if (mode & 2) kick_body(S, b, dt / 2.);
  if (mode & 8) kick_body(S, b, dt / 2.);
  if (mode & 4) { drift_body(S, b, dt); check_displacement(S, b, skin); }
}

__global__ void k_rk2(RbxScene S, int stage, double dt, int fix_q7, double skin) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= S.n_bodies) return;
  const int i3 = 3 * b, i9 = 9 * b;
  if (stage == 0) {  // py_initialize :407-419
#pragma unroll
    for (int j = 0; j < 3; j++) {
      S.xcm0[i3 + j] = S.xcm[i3 + j];
      S.vcm0[i3 + j] = S.vcm[i3 + j];
      // :415 writes ang_mom0[j], j < 3: the first body of the array (Q7)
      if (fix_q7 || (S.body_first ? S.body_first[b] == b : b == 0))
        S.ang_mom0[i3 + j] = S.ang_mom[i3 + j];
    }
#pragma unroll
    for (int k = 0; k < 9; k++) S.R0[i9 + k] = S.R[i9 + k];
    return;
  }
  const double h = (stage == 1) ? dt / 2. : dt;
  const double M = S.total_mass[b];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    S.xcm[i3 + j] = S.xcm0[i3 + j] + h * S.vcm[i3 + j];
    S.vcm[i3 + j] = S.vcm0[i3 + j] + h * S.force[i3 + j] / M;
  }
  double om[3] = {S.omega[i3], S.omega[i3 + 1], S.omega[i3 + 2]};
  rotate_body(S.R + i9, S.R0 + i9, om, h, S.iinv_b + i9, S.iinv_g + i9);
  double L[3];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    L[j] = S.ang_mom0[i3 + j] + (h * S.torque[i3 + j]);
    S.ang_mom[i3 + j] = L[j];
  }
  matvec3(S.iinv_g + i9, L, om);
#pragma unroll
  for (int j = 0; j < 3; j++) S.omega[i3 + j] = om[j];
  check_displacement(S, b, skin);
}

__global__ void __launch_bounds__(256) k_pose(RbxScene S, int flags) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= S.n_rigid) return;
  const int b = S.body[p];
  const int i9 = 9 * b, i3 = 3 * b;
  const double x0 = S.dx0[p], y0 = S.dy0[p], z0 = S.dz0[p];
  if (flags & RBX_POSE_VEL) {
    const double *Rv = ((flags & RBX_POSE_VEL_PREV) ? S.R_prev : S.R) + i9;
    double u, v, w;
    rbx_point_velocity(Rv, S.omega + i3, S.vcm + i3, x0, y0, z0, u, v, w);
    S.u[p] = u;
    S.v[p] = v;
    S.w[p] = w;
  }
  if (flags & (RBX_POSE_POS | RBX_POSE_NORMALS)) {
    double R[9];
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = S.R[i9 + k];
    if (flags & RBX_POSE_POS) {
      const double dx = (R[0] * x0 + R[1] * y0 + R[2] * z0);
      const double dy = (R[3] * x0 + R[4] * y0 + R[5] * z0);
      const double dz = (R[6] * x0 + R[7] * y0 + R[8] * z0);
      const double xn = S.xcm[i3] + dx, yn = S.xcm[i3 + 1] + dy, zn = S.xcm[i3 + 2] + dz;
      S.x[p] = xn;
      S.y[p] = yn;
      S.z[p] = zn;
      // FP32 copy for the first pass of the contact evaluation (k_filter)
      if (S.pos32) {
        const float hf = S.h_uniform > 0. ? (float)S.h_uniform : (float)S.h[p];
        reinterpret_cast<float4 *>(S.pos32)[p] =
            make_float4((float)(xn - S.origin[0]), (float)(yn - S.origin[1]),
                        (float)(zn - S.origin[2]), hf);
      }
    }
    if ((flags & RBX_POSE_NORMALS) && S.normal && S.is_boundary && S.is_boundary[p] == 1) {
      const double n0 = S.normal0[3 * p], n1 = S.normal0[3 * p + 1], n2 = S.normal0[3 * p + 2];
      S.normal[3 * p] = (R[0] * n0 + R[1] * n1 + R[2] * n2);
      S.normal[3 * p + 1] = (R[3] * n0 + R[4] * n1 + R[5] * n2);
      S.normal[3 * p + 2] = (R[6] * n0 + R[7] * n1 + R[8] * n2);
    }
  }
}

// halo payload rows {x, y, z, u, v, w, h, dem_id}: thread <-> (particle, column)
__global__ void k_halo_pack(RbxScene S, const int64_t *index, int n, double *rows,
                            int body_vel) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n * 8) return;
  const int c = k & 7;
  const long long q = index[k >> 3];
  if (body_vel && c >= 3 && c < 6 && q < S.n_rigid) {
    // stage-1 velocity of a body particle, formed here: u, v, w are not kept
    // current inside a step (RBX_PARAM_BODY_VEL)
    const int b = S.body[q];
    double uvw[3];
    rbx_point_velocity(S.R_prev + 9 * b, S.omega + 3 * b, S.vcm + 3 * b, S.dx0[q], S.dy0[q],
                       S.dz0[q], uvw[0], uvw[1], uvw[2]);
    rows[k] = uvw[c - 3];
    return;
  }
  const double *col[7] = {S.x, S.y, S.z, S.u, S.v, S.w, S.h};
  rows[k] = c < 7 ? col[c][q] : (double)S.dem_id[q];
}
// a static particle (wall, halo) that is now further than skin / 2 from where
// it was when the neighbour lists were built asks for a rebuild
__device__ __forceinline__ void check_static(const RbxScene &S, int q, double x, double y,
                                             double z, double skin) {
  if (!S.static_ref || !S.rebuild) return;
  const double *r = S.static_ref + 3 * (size_t)(q - S.n_rigid);
  const double dx = x - r[0], dy = y - r[1], dz = z - r[2];
  if (!(dx * dx + dy * dy + dz * dz <= 0.25 * skin * skin)) atomicOr(S.rebuild, 1u);
}

__global__ void k_halo_unpack(RbxScene S, int first, int n, const double *rows, double skin) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n * 8) return;
  const int c = k & 7;
  const int q = first + (k >> 3);
  double *col[7] = {S.x, S.y, S.z, S.u, S.v, S.w, const_cast<double *>(S.h)};
  if (c < 7) col[c][q] = rows[k];
  else const_cast<int32_t *>(S.dem_id)[q] = (int32_t)rows[k];
  if (c == 0) {
    const double *r = rows + (size_t)(k >> 3) * 8;
    if (S.pos32)
      reinterpret_cast<float4 *>(S.pos32)[q] =
          make_float4((float)(r[0] - S.origin[0]), (float)(r[1] - S.origin[1]),
                      (float)(r[2] - S.origin[2]), (float)r[6]);
    check_static(S, q, r[0], r[1], r[2], skin);
  }
}

__global__ void k_static_update(RbxScene S, int first, int n, const double *x, const double *y,
                                const double *z, const double *u, const double *v,
                                const double *w, double skin) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int q = first + k;
  if (x) S.x[q] = x[k];
  if (y) S.y[q] = y[k];
  if (z) S.z[q] = z[k];
  if (u) S.u[q] = u[k];
  if (v) S.v[q] = v[k];
  if (w) S.w[q] = w[k];
  if (x || y || z) {
    const double xn = S.x[q], yn = S.y[q], zn = S.z[q];
    if (S.pos32)
      reinterpret_cast<float4 *>(S.pos32)[q] =
          make_float4((float)(xn - S.origin[0]), (float)(yn - S.origin[1]),
                      (float)(zn - S.origin[2]), (float)S.h[q]);
    check_static(S, q, xn, yn, zn, skin);
  }
}

int launch_bodies(const RbxScene *S, int mode, double dt, double skin, cudaStream_t st) {
  if (S->n_bodies <= 0) return RBX_OK;
  const int T = 128;
  if (mode & 1) k_reduce<<<rbx_blocks((long long)S->n_bodies * 32, T), T, 0, st>>>(*S);
  if (mode & ~1) k_bodies<<<rbx_blocks(S->n_bodies, T), T, 0, st>>>(*S, mode & ~1, dt, skin);
  RBX_CHECK_LAUNCH();
  return RBX_OK;
}

}  // namespace

extern "C" int rbx_reduce_bodies(const RbxScene *scene, void *stream) {
  if (!scene) return RBX_ERR_INVALID;
  return launch_bodies(scene, 1, 0., 0., (cudaStream_t)stream);
}

extern "C" int rbx_gtvf_kick(const RbxScene *scene, double dt, void *stream) {
  if (!scene) return RBX_ERR_INVALID;
  return launch_bodies(scene, 2, dt, 0., (cudaStream_t)stream);
}

extern "C" int rbx_gtvf_drift(const RbxScene *scene, double dt, double skin, void *stream) {
  if (!scene) return RBX_ERR_INVALID;
  return launch_bodies(scene, 4, dt, skin, (cudaStream_t)stream);
}

extern "C" int rbx_pose_particles(const RbxScene *scene, int flags, void *stream) {
  if (!scene) return RBX_ERR_INVALID;
  if (scene->n_rigid <= 0) return RBX_OK;
  if ((flags & RBX_POSE_VEL_PREV) && !scene->R_prev) return RBX_ERR_INVALID;
  k_pose<<<rbx_blocks(scene->n_rigid, 256), 256, 0, (cudaStream_t)stream>>>(*scene, flags);
  RBX_CHECK_LAUNCH();
  return RBX_OK;
}

extern "C" int rbx_halo_pack(const RbxScene *scene, const int64_t *index, int32_t n,
                             double *rows, int body_vel, void *stream) {
  if (!scene || n < 0 || (n > 0 && (!index || !rows))) return RBX_ERR_INVALID;
  if (body_vel && !scene->R_prev) return RBX_ERR_INVALID;
  if (n == 0) return RBX_OK;
  k_halo_pack<<<rbx_blocks((long long)n * 8, 256), 256, 0, (cudaStream_t)stream>>>(*scene, index, n, rows, body_vel);
  RBX_CHECK_LAUNCH();
  return RBX_OK;
}

extern "C" int rbx_static_update(const RbxScene *scene, int32_t first, int32_t n,
                                 const double *x, const double *y, const double *z,
                                 const double *u, const double *v, const double *w,
                                 double skin, void *stream) {
  if (!scene || n < 0 || first < scene->n_rigid || (long long)first + n > scene->n_total)
    return RBX_ERR_INVALID;
  if (n == 0) return RBX_OK;
  k_static_update<<<rbx_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(*scene, first, n, x, y, z,
                                                                        u, v, w, skin);
  RBX_CHECK_LAUNCH();
  return RBX_OK;
}

extern "C" int rbx_halo_unpack(const RbxScene *scene, int32_t first, int32_t n,
                               const double *rows, double skin, void *stream) {
  if (!scene || n < 0 || first < 0 || (long long)first + n > scene->n_total ||
      (n > 0 && !rows)) return RBX_ERR_INVALID;
  if (n == 0) return RBX_OK;
  k_halo_unpack<<<rbx_blocks((long long)n * 8, 256), 256, 0, (cudaStream_t)stream>>>(*scene, first, n, rows, skin);
  RBX_CHECK_LAUNCH();
  return RBX_OK;
}

extern "C" int rbx_rk2_stage(const RbxScene *scene, int stage, double dt, int fix_q7,
                             double skin, void *stream) {
  if (!scene || stage < 0 || stage > 2) return RBX_ERR_INVALID;
  if (!scene->xcm0 || !scene->vcm0 || !scene->ang_mom0 || !scene->R0) return RBX_ERR_INVALID;
  if (scene->n_bodies <= 0) return RBX_OK;
  k_rk2<<<rbx_blocks(scene->n_bodies, 128), 128, 0, (cudaStream_t)stream>>>(*scene, stage, dt, fix_q7, skin);
  RBX_CHECK_LAUNCH();
  return RBX_OK;
}

// Whole GTVF step: kick + drift (one body launch), positions + stage-1
// velocities (one particle launch), cell list, contact, reduce + kick (one
// body launch), stage-3 velocities.
// ---- list rebuild, conditional inside a captured graph -----------------------
__global__ void k_set_conditional(cudaGraphConditionalHandle h, const uint32_t *flag) {
  cudaGraphSetConditional(h, *flag != 0u ? 1u : 0u);
}

static int rebuild_lists_plain(const RbxScene *scene, const RbxPoints *src, const RbxCells *cells,
                               const RbxParams *par, void *ws, size_t ws_bytes, cudaStream_t st) {
  int rc = rbx_cells_build(src, cells, par->reach, scene->status, ws, ws_bytes, st);
  if (rc) return rc;
  return rbx_contact_neighbours(scene, cells, par, st);
}

// Cell list + neighbour lists.  Every kernel of the rebuild returns at once
// while *rebuild == 0, but 13 launches that do nothing still cost 80 us of a
// 1.8 ms step.  When `st` is being captured and the caller gave a second
// stream, the rebuild is captured into the body of a conditional IF node on
// *rebuild instead (CUDA 12.4 graph conditionals): steps that reuse the lists
// skip it on the device, without a host decision.
static int rebuild_lists(const RbxScene *scene, const RbxPoints *src, const RbxCells *cells,
                         const RbxParams *par, void *ws, size_t ws_bytes, cudaStream_t st) {
  cudaStream_t aux = (cudaStream_t)scene->aux_stream;
  if (aux && scene->rebuild && cells->cond == scene->rebuild && par->skin > 0.) {
    cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
    cudaGraph_t graph = nullptr;
    const cudaGraphNode_t *deps = nullptr;
    size_t ndeps = 0;
    unsigned long long id = 0;
    const cudaError_t e1 = cudaStreamGetCaptureInfo_v2(st, &status, &id, &graph, &deps, &ndeps);
    const bool dbg = getenv("RBX_DEBUG_GRAPH") != nullptr;
    if (dbg) fprintf(stderr, "[rbx] capture info: err %d status %d graph %p deps %zu\n", (int)e1,
                     (int)status, (void *)graph, ndeps);
    if (e1 == cudaSuccess && status == cudaStreamCaptureStatusActive && graph) {
      cudaGraphConditionalHandle handle;
      const cudaError_t e2 =
          cudaGraphConditionalHandleCreate(&handle, graph, 0, cudaGraphCondAssignDefault);
      if (dbg) fprintf(stderr, "[rbx] conditional handle: err %d (%s)\n", (int)e2,
                       cudaGetErrorString(e2));
      if (e2 == cudaSuccess) {
        k_set_conditional<<<1, 1, 0, st>>>(handle, scene->rebuild);
        if (cudaStreamGetCaptureInfo_v2(st, &status, &id, &graph, &deps, &ndeps) != cudaSuccess)
          return RBX_ERR_LAUNCH;
        cudaGraphNodeParams np = {};
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = handle;
        np.conditional.type = cudaGraphCondTypeIf;
        np.conditional.size = 1;
        cudaGraphNode_t node;
        if (cudaGraphAddNode(&node, graph, deps, ndeps, &np) != cudaSuccess) return RBX_ERR_LAUNCH;
        cudaGraph_t body = np.conditional.phGraph_out[0];
        if (cudaStreamBeginCaptureToGraph(aux, body, nullptr, nullptr, 0,
                                          cudaStreamCaptureModeRelaxed) != cudaSuccess)
          return RBX_ERR_LAUNCH;
        const int rc = rebuild_lists_plain(scene, src, cells, par, ws, ws_bytes, aux);
        cudaGraph_t got = nullptr;
        if (cudaStreamEndCapture(aux, &got) != cudaSuccess || rc) return rc ? rc : RBX_ERR_LAUNCH;
        if (cudaStreamUpdateCaptureDependencies(st, &node, 1, cudaStreamSetCaptureDependencies) !=
            cudaSuccess)
          return RBX_ERR_LAUNCH;
        return RBX_OK;
      }
      cudaGetLastError();      // no conditional nodes here: plain launches
    } else {
      cudaGetLastError();
    }
  }
  return rebuild_lists_plain(scene, src, cells, par, ws, ws_bytes, st);
}

extern "C" int rbx_gtvf_step(const RbxScene *scene, const RbxPoints *src, const RbxCells *cells,
                             const RbxParams *params, void *workspace, size_t workspace_bytes,
                             int flags, void *stream) {
  if (!scene || !src || !cells || !params) return RBX_ERR_INVALID;
  if (!scene->R_prev) return RBX_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const bool pre = !(flags & 4), post = !(flags & 2);
  int rc;
  if (pre) {
    if ((rc = launch_bodies(scene, 8 | 4, params->dt, params->skin, st))) return rc;
    // Positions only.  The stage-1 velocities (post-kick omega, pre-drift R)
    // are needed for the ~1 % of the particles in contact, which the contact
    // law forms itself (RBX_PARAM_BODY_VEL); u, v, w and the rotated normals
    // are written once, at the end of the step.
    if ((rc = rbx_pose_particles(scene, RBX_POSE_POS, stream))) return rc;
  }
  if (post) {
    RbxParams par = *params;
    par.flags |= RBX_PARAM_BODY_VEL;
    if ((rc = rebuild_lists(scene, src, cells, &par, workspace, workspace_bytes, st))) return rc;
    if ((rc = rbx_contact_slots(scene, cells, &par, nullptr, stream))) return rc;
    // reduce (k_reduce, warp per body) and kick (k_bodies, thread per body)
    // are two launches: with the kick's divisions on lane 0 of every reduce
    // warp, the warp sat on its slot three times longer than its loads take
    if ((rc = launch_bodies(scene, 1 | 2, params->dt, 0., st))) return rc;
    if (!(flags & 1))
      if ((rc = rbx_pose_particles(scene, RBX_POSE_VEL | RBX_POSE_NORMALS, stream))) return rc;
  }
  return RBX_OK;
}

extern "C" int rbx_version(void) { return RBX_VERSION; }

extern "C" const char *rbx_strerror(int code) {
  switch (code) {
    case RBX_OK: return "ok";
    case RBX_ERR_INVALID: return "invalid argument";
    case RBX_ERR_WORKSPACE: return "workspace too small";
    case RBX_ERR_LAUNCH: return "CUDA launch failed";
    case RBX_ERR_NO_DEVICE: return "no CUDA device";
    default: return "unknown error";
  }
}

// sizeof of the ABI structs, so that bindings can verify their mirrors
extern "C" size_t rbx_sizeof(int which) {
  switch (which) {
    case 0: return sizeof(RbxGridInfo);
    case 1: return sizeof(RbxPoints);
    case 2: return sizeof(RbxCells);
    case 3: return sizeof(RbxScene);
    case 4: return sizeof(RbxParams);
    case 5: return sizeof(RbxDiag);
    case 6: return sizeof(RbxDemScene);
    case 7: return sizeof(RbxCanelas);
    default: return 0;
  }
}
