// Setup path on the device: boundary-particle
// identification of one particle array,
//   ComputeNormals -> SmoothNormals   [upstream pysph.sph.isph.wall_normal;
//                                      text: boundary_particles.py:71-135]
//   IdentifyBoundaryParticleCosAngle  boundary_particles.py:22-68
// as three thread-per-particle passes over the cell list of that array (all
// particles binned).  One-off work: the host version (compat/sph_evaluator.py,
// KD-tree + NumPy) takes minutes at 10^7 particles.
#include "rbx_common.cuh"

namespace {

// gradient factor of the quintic spline: dW/dr / r  (QuinticSpline.gradient)
template <int DIM>
__device__ __forceinline__ double quintic_dwdr_over_r(double rij, double h) {
  const double M_1_PI_ = 0.31830988618379067154;
  if (!(rij > 1e-12)) return 0.;
  const double h1 = 1. / h;
  const double q = rij * h1;
  double fac;
  if (DIM == 2) fac = (M_1_PI_ * 7.0 / 478.0) * h1 * h1;
  else if (DIM == 3) fac = (M_1_PI_ / 120.0) * h1 * h1 * h1;
  else fac = (1.0 / 120.0) * h1;
  const double t3 = 3. - q, t2 = 2. - q, t1 = 1. - q;
  double val;
  if (q > 3.0) val = 0.0;
  else {
    val = -5.0 * t3 * t3 * t3 * t3;
    if (q <= 2.0) val += 30.0 * t2 * t2 * t2 * t2;
    if (q <= 1.0) val -= 75.0 * t1 * t1 * t1 * t1;
  }
  return val * fac * h1 / rij;
}

struct Stencil {
  int cx0, cx1, cy0, cy1, cz0, cz1;
};

__device__ __forceinline__ Stencil stencil_of(const RbxGridInfo &gi, double px, double py,
                                              double pz) {
  const double reach = gi.cell * (1.0 + 1e-9);
  Stencil s;
  s.cx0 = rbx_cell_coord(px - reach, gi.x0, gi.inv_cell, gi.nx);
  s.cx1 = rbx_cell_coord(px + reach, gi.x0, gi.inv_cell, gi.nx);
  s.cy0 = rbx_cell_coord(py - reach, gi.y0, gi.inv_cell, gi.ny);
  s.cy1 = rbx_cell_coord(py + reach, gi.y0, gi.inv_cell, gi.ny);
  s.cz0 = rbx_cell_coord(pz - reach, gi.z0, gi.inv_cell, gi.nz);
  s.cz1 = rbx_cell_coord(pz + reach, gi.z0, gi.inv_cell, gi.nz);
  return s;
}

// pass: 0 ComputeNormals, 1 SmoothNormals, 2 IdentifyBoundaryParticleCosAngle
template <int DIM, int PASS>
__global__ void k_boundary(RbxPoints D, RbxCells C, double rs, const double *m, const double *rho,
                           double *normal_tmp, double *normal, int32_t *is_boundary) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= D.n) return;
  const RbxGridInfo gi = *C.info;
  const int g = D.index ? D.index[k] : k;
  const double px = D.x[g], py = D.y[g], pz = D.z[g], ph = D.h[g];
  const double rs2 = rs * rs;
  const double hi2 = rbx_h2(rs2, ph);
  const Stencil st = stencil_of(gi, px, py, pz);
  double a0 = 0., a1 = 0., a2 = 0.;
  double n0 = 0., n1 = 0., n2 = 0.;
  bool candidate = false, killed = false;
  if (PASS == 1) { a0 = normal[3 * g]; a1 = normal[3 * g + 1]; a2 = normal[3 * g + 2]; }
  if (PASS == 2) {
    n0 = normal[3 * g]; n1 = normal[3 * g + 1]; n2 = normal[3 * g + 2];
    candidate = (n0 * n0 + n1 * n1 + n2 * n2) > 1e-6;
  }
  if (PASS != 2 || candidate) {
    for (int cz = st.cz0; cz <= st.cz1 && !killed; cz++)
      for (int cy = st.cy0; cy <= st.cy1 && !killed; cy++) {
        const int r0 = (cz * gi.ny + cy) * gi.nx;
        const int s = C.cell_start[r0 + st.cx0], e = C.cell_start[r0 + st.cx1 + 1];
        for (int q = s; q < e; q++) {
          const double x0 = px - C.sx[q], x1 = py - C.sy[q], x2 = pz - C.sz[q];
          const double r2 = rbx_r2(x0, x1, x2);
          const double hj = C.sh[q];
          if (!(r2 < hi2 || r2 < rbx_h2(rs2, hj))) continue;   // NNPS
          const double rij = sqrt(r2);
          const int j = C.gidx[q];
          if (PASS == 0) {
            const double fac = -m[j] / rho[j];
            const double gw = quintic_dwdr_over_r<DIM>(rij, 0.5 * (ph + hj));
            a0 += fac * (gw * x0); a1 += fac * (gw * x1); a2 += fac * (gw * x2);
          } else if (PASS == 1) {
            const double fac = m[j] / rho[j] * rbx_quintic<DIM>(rij, 0.5 * (ph + hj));
            a0 += fac * normal_tmp[3 * j]; a1 += fac * normal_tmp[3 * j + 1];
            a2 += fac * normal_tmp[3 * j + 2];
          } else {
            if (rij > 1e-9 * ph && rij < 2. * ph) {
              const double dot = -(n0 * x0 + n1 * x1 + n2 * x2);
              if (dot / rij > 0.5) { killed = true; break; }
            }
          }
        }
      }
  }
  if (PASS == 0) {
    const double mag = sqrt(a0 * a0 + a1 * a1 + a2 * a2);
    const bool ok = mag > 0.25 / ph;
    normal_tmp[3 * g] = ok ? a0 / mag : 0.;
    normal_tmp[3 * g + 1] = ok ? a1 / mag : 0.;
    normal_tmp[3 * g + 2] = ok ? a2 / mag : 0.;
    normal[3 * g] = 0.; normal[3 * g + 1] = 0.; normal[3 * g + 2] = 0.;
  } else if (PASS == 1) {
    const double mag = sqrt(a0 * a0 + a1 * a1 + a2 * a2);
    const bool ok = mag > 1e-3;
    normal[3 * g] = ok ? a0 / mag : 0.;
    normal[3 * g + 1] = ok ? a1 / mag : 0.;
    normal[3 * g + 2] = ok ? a2 / mag : 0.;
  } else {
    is_boundary[g] = (candidate && !killed) ? 1 : 0;
  }
}

template <int DIM>
int launch_all(const RbxPoints *pts, const RbxCells *cells, double rs, const double *m,
               const double *rho, double *normal_tmp, double *normal, int32_t *isb,
               cudaStream_t st) {
  const int T = 128, nb = rbx_blocks(pts->n, T);
  k_boundary<DIM, 0><<<nb, T, 0, st>>>(*pts, *cells, rs, m, rho, normal_tmp, normal, isb);
  k_boundary<DIM, 1><<<nb, T, 0, st>>>(*pts, *cells, rs, m, rho, normal_tmp, normal, isb);
  k_boundary<DIM, 2><<<nb, T, 0, st>>>(*pts, *cells, rs, m, rho, normal_tmp, normal, isb);
  RBX_CHECK_LAUNCH();
  return RBX_OK;
}

// set_total_mass, set_center_of_mass, set_moment_of_inertia_izz,
// set_moment_of_inertia_and_its_inverse, set_body_frame_position_vectors
// (rigid_body_common.py:21-107): one warp per body over its particles
// [start[b], start[b + 1]), two passes (mass and centre of mass, then the
// tensor about it), fixed shuffle trees; the 3x3 inverse by cofactors on lane
// 0 (the reference calls np.linalg.inv: LU, same result to rounding; a
// singular tensor gives inf / nan there and here).
__global__ void __launch_bounds__(128)
k_setup_bodies(int nb, const int32_t *start, const double *x, const double *y, const double *z,
               const double *m, double *total_mass, double *xcm, double *izz, double *I_out,
               double *Iinv_out, double *dx0, double *dy0, double *dz0) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= nb) return;
  const int q0 = start[b], q1 = start[b + 1];
  double M = 0., sx = 0., sy = 0., sz = 0.;
  for (int q = q0 + lane; q < q1; q += 32) {
    const double mq = m[q];
    M += mq; sx += mq * x[q]; sy += mq * y[q]; sz += mq * z[q];
  }
  M = rbx_warp_sum(M); sx = rbx_warp_sum(sx); sy = rbx_warp_sum(sy); sz = rbx_warp_sum(sz);
  const double cx = sx / M, cy = sy / M, cz = sz / M;
  double ixx = 0., iyy = 0., izz_ = 0., ixy = 0., ixz = 0., iyz = 0.;
  for (int q = q0 + lane; q < q1; q += 32) {
    const double mq = m[q];
    const double dx = x[q] - cx, dy = y[q] - cy, dz = z[q] - cz;
    dx0[q] = dx; dy0[q] = dy; dz0[q] = dz;
    ixx += mq * (dy * dy + dz * dz);
    iyy += mq * (dx * dx + dz * dz);
    izz_ += mq * (dx * dx + dy * dy);
    ixy += mq * dx * dy; ixz += mq * dx * dz; iyz += mq * dy * dz;
  }
  ixx = rbx_warp_sum(ixx); iyy = rbx_warp_sum(iyy); izz_ = rbx_warp_sum(izz_);
  ixy = rbx_warp_sum(ixy); ixz = rbx_warp_sum(ixz); iyz = rbx_warp_sum(iyz);
  if (lane == 0) {
    total_mass[b] = M;
    xcm[3 * b] = cx; xcm[3 * b + 1] = cy; xcm[3 * b + 2] = cz;
    if (izz) izz[b] = izz_;
    const double I[9] = {ixx, -ixy, -ixz, -ixy, iyy, -iyz, -ixz, -iyz, izz_};
    for (int k = 0; k < 9; k++) I_out[9 * b + k] = I[k];
    const double c00 = I[4] * I[8] - I[5] * I[7], c01 = I[5] * I[6] - I[3] * I[8],
                 c02 = I[3] * I[7] - I[4] * I[6];
    const double det = I[0] * c00 + I[1] * c01 + I[2] * c02;
    const double inv[9] = {c00 / det, (I[2] * I[7] - I[1] * I[8]) / det,
                           (I[1] * I[5] - I[2] * I[4]) / det,
                           c01 / det, (I[0] * I[8] - I[2] * I[6]) / det,
                           (I[2] * I[3] - I[0] * I[5]) / det,
                           c02 / det, (I[1] * I[6] - I[0] * I[7]) / det,
                           (I[0] * I[4] - I[1] * I[3]) / det};
    for (int k = 0; k < 9; k++) Iinv_out[9 * b + k] = inv[k];
  }
}

}  // namespace

extern "C" int rbx_setup_bodies(int32_t n_bodies, const int32_t *body_start, const double *x,
                                const double *y, const double *z, const double *m,
                                double *total_mass, double *xcm, double *izz, double *inertia,
                                double *inertia_inverse, double *dx0, double *dy0, double *dz0,
                                void *stream) {
  if (n_bodies < 0 || !body_start || !x || !y || !z || !m || !total_mass || !xcm || !inertia ||
      !inertia_inverse || !dx0 || !dy0 || !dz0)
    return RBX_ERR_INVALID;
  if (n_bodies == 0) return RBX_OK;
  k_setup_bodies<<<rbx_blocks((long long)n_bodies * 32, 128), 128, 0, (cudaStream_t)stream>>>(
      n_bodies, body_start, x, y, z, m, total_mass, xcm, izz, inertia, inertia_inverse, dx0, dy0,
      dz0);
  RBX_CHECK_LAUNCH();
  return RBX_OK;
}

extern "C" int rbx_boundary_identify(const RbxPoints *pts, const RbxCells *cells, int dim,
                                     double radius_scale, const double *m, const double *rho,
                                     double *normal_tmp, double *normal, int32_t *is_boundary,
                                     void *stream) {
  if (!pts || !cells || !m || !rho || !normal_tmp || !normal || !is_boundary)
    return RBX_ERR_INVALID;
  if (pts->n <= 0) return RBX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dim == 3) return launch_all<3>(pts, cells, radius_scale, m, rho, normal_tmp, normal, is_boundary, st);
  if (dim == 2) return launch_all<2>(pts, cells, radius_scale, m, rho, normal_tmp, normal, is_boundary, st);
  return RBX_ERR_INVALID;
}
