// DEMScheme path: linear visco-elastic contact with Coulomb friction and a
// per-pair tangential displacement history (Luding 2008), plus the
// per-particle velocity-Verlet stepper.
//
//   UpdateTangentialContactsLVCDisplacement.initialize_pair  dem.py:208-293
//   BodyForce.initialize                     rigid_body_common.py:115-125
//   LVCDisplacement.loop                                     dem.py:35-205
//   DEMStep.stage1/2/3                                       dem.py:595-625
//
// One thread per destination (granular) particle; the pair loop walks the
// cell list built over ALL particles (every sphere is a source).  The history
// list of a particle (`limit` entries: source-local index, source dem_id,
// tangential displacement) is searched linearly, exactly as the reference
// does.  HBM-bound streaming plus a short gather per neighbour; the scheme is
// not instantiated by any script of the reference, so this kernel is kept
// simple rather than tuned.
#include "rbx_common.cuh"

namespace {

__global__ void k_dem_update(RbxDemScene S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.n_dest) return;
  const int limit = S.limit;
  const int p = i * limit;
  const double xi = S.x[i], yi = S.y[i], zi = S.z[i], ri = S.rad_s[i];
  for (int a = 0; a < S.n_arrays; a++) {          // sources in scheme order
    const int a0 = S.arr_start[a], an = S.arr_start[a + 1] - a0;
    const int total = S.total_tng[i];
    int last = p + total - 1;
    int k = p, count = 0;
    while (count < total) {
      const int sidx = S.tng_idx[k];
      const int dem = S.tng_dem[k];
      if (sidx == -1) break;
      // quirk Q13: the reference indexes THIS source array with an index that
      // may belong to another one; out-of-range reads are skipped here
      if (sidx < an && dem == S.dem_id[a0 + sidx]) {
        const int g = a0 + sidx;
        const double x0 = xi - S.x[g], x1 = yi - S.y[g], x2 = zi - S.z[g];
        const double rij = sqrt(x0 * x0 + x1 * x1 + x2 * x2);
        const double overlap = ri + S.rad_s[g] - rij;
        if (overlap <= 0.) {
          if (k == last) {
            S.tng_idx[k] = -1; S.tng_dem[k] = -1;
            S.tng_x[k] = 0.; S.tng_y[k] = 0.; S.tng_z[k] = 0.;
          } else {
            S.tng_idx[k] = S.tng_idx[last]; S.tng_idx[last] = -1;
            S.tng_x[k] = S.tng_x[last]; S.tng_x[last] = 0.;
            S.tng_y[k] = S.tng_y[last]; S.tng_y[last] = 0.;
            S.tng_z[k] = S.tng_z[last]; S.tng_z[last] = 0.;
            S.tng_dem[k] = S.tng_dem[last]; S.tng_dem[last] = -1;
            last -= 1;
          }
          S.total_tng[i] -= 1;
        } else {
          k = k + 1;
        }
      } else {
        k = k + 1;
      }
      count += 1;
    }
  }
}

__global__ void k_dem_force(RbxDemScene S, RbxCells C, RbxParams P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.n_dest) return;
  const RbxGridInfo gi = *C.info;
  const double px = S.x[i], py = S.y[i], pz = S.z[i], ph = S.h[i];
  const double ud = S.u[i], vd = S.v[i], wd = S.w[i];
  const double dwx = S.wx[i], dwy = S.wy[i], dwz = S.wz[i];
  const double md = S.m[i], rd = S.rad_s[i];
  const int row = S.tbl_row[i];
  const int limit = S.limit;
  const int p = i * limit;
  double fx = md * P.gx, fy = md * P.gy, fz = md * P.gz;   // BodyForce
  double tx = S.torx[i], ty = S.tory[i], tz = S.torz[i];   // never reset
  int tot = S.total_tng[i];
  unsigned st = 0u;

  const double rs2 = P.radius_scale * P.radius_scale;
  const double hi2 = rbx_h2(rs2, ph);
  const double reach = gi.cell * (1.0 + 1e-9);
  const int cx0 = rbx_cell_coord(px - reach, gi.x0, gi.inv_cell, gi.nx);
  const int cx1 = rbx_cell_coord(px + reach, gi.x0, gi.inv_cell, gi.nx);
  const int cy0 = rbx_cell_coord(py - reach, gi.y0, gi.inv_cell, gi.ny);
  const int cy1 = rbx_cell_coord(py + reach, gi.y0, gi.inv_cell, gi.ny);
  const int cz0 = rbx_cell_coord(pz - reach, gi.z0, gi.inv_cell, gi.nz);
  const int cz1 = rbx_cell_coord(pz + reach, gi.z0, gi.inv_cell, gi.nz);
  for (int cz = cz0; cz <= cz1; cz++)
    for (int cy = cy0; cy <= cy1; cy++) {
      const int r0 = (cz * gi.ny + cy) * gi.nx;
      const int s = C.cell_start[r0 + cx0], e = C.cell_start[r0 + cx1 + 1];
      for (int q = s; q < e; q++) {
        const double x0 = px - C.sx[q], x1 = py - C.sy[q], x2 = pz - C.sz[q];
        const double r2 = rbx_r2(x0, x1, x2);
        if (!(r2 < hi2 || r2 < rbx_h2(rs2, C.sh[q]))) continue;   // NNPS
        const double rij = sqrt(r2);
        const int g = C.gidx[q];
        double overlap = -1.;
        if (rij > 0) overlap = rd + S.rad_s[g] - rij;
        if (!(overlap > 0)) continue;
        const int sdem = C.sdem[q];
        const int sloc = g - S.arr_off[g];                // source-local idx
        const double rinv = 1.0 / rij;
        const double nx = x0 * rinv, ny = x1 * rinv, nz = x2 * rinv;
        const double a_i = rd - overlap / 2.;
        const double a_j = S.rad_s[g] - overlap / 2.;
        const double swx = S.wx[g], swy = S.wy[g], swz = S.wz[g];
        const double vi_x = ud + (dwy * nz - dwz * ny) * a_i;
        const double vi_y = vd + (dwz * nx - dwx * nz) * a_i;
        const double vi_z = wd + (dwx * ny - dwy * nx) * a_i;
        const double vj_x = S.u[g] + (-swy * nz + swz * ny) * a_j;
        const double vj_y = S.v[g] + (-swz * nx + swx * nz) * a_j;
        const double vj_z = S.w[g] + (-swx * ny + swy * nx) * a_j;
        const double vij_x = vi_x - vj_x, vij_y = vi_y - vj_y, vij_z = vi_z - vj_z;
        const double vn = vij_x * nx + vij_y * ny + vij_z * nz;
        const double vt_x = vij_x - vn * nx, vt_y = vij_y - vn * ny, vt_z = vij_z - vn * nz;
        const double ms = S.m[g];
        const double m_eff = md * ms / (md + ms);
        const double eta_n = S.alpha[row + sdem] * sqrt(m_eff);
        const double fn = S.kn[row + sdem] * overlap + eta_n * -vn;
        int found_at = -1;
        for (int k = p; k < p + tot; k++)
          if (sloc == S.tng_idx[k] && sdem == S.tng_dem[k]) { found_at = k; break; }
        double ft_x = 0., ft_y = 0., ft_z = 0.;
        if (found_at < 0) {
          if (tot < limit) {             // new contact: no spring force yet
            S.tng_idx[p + tot] = sloc;
            S.tng_dem[p + tot] = sdem;
            tot++;
          } else {
            st |= RBX_STATUS_LVC_OVERFLOW;   // the reference writes past it (Q14)
          }
        } else {
          double gx_ = S.tng_x[found_at], gy_ = S.tng_y[found_at], gz_ = S.tng_z[found_at];
          const double tdn = (gx_ * nx + gy_ * ny + gz_ * nz);
          gx_ = gx_ - tdn * nx; gy_ = gy_ - tdn * ny; gz_ = gz_ - tdn * nz;
          const double kt = S.kt[row + sdem];
          const double kt_1 = 1. / kt;
          ft_x = -kt * gx_ - eta_n * vt_x;
          ft_y = -kt * gy_ - eta_n * vt_y;
          ft_z = -kt * gz_ - eta_n * vt_z;
          const double ft_magn = sqrt(ft_x * ft_x + ft_y * ft_y + ft_z * ft_z);
          double ux = 0., uy = 0., uz = 0.;
          if (ft_magn > 1e-12) { ux = ft_x / ft_magn; uy = ft_y / ft_magn; uz = ft_z / ft_magn; }
          const double fn_mu = S.mu[row + sdem] * fn;
          if (ft_magn > fn_mu) {
            ft_x = fn_mu * ux; ft_y = fn_mu * uy; ft_z = fn_mu * uz;
            gx_ = -kt_1 * (fn_mu * ux + eta_n * vt_x);
            gy_ = -kt_1 * (fn_mu * uy + eta_n * vt_y);
            gz_ = -kt_1 * (fn_mu * uz + eta_n * vt_z);
          } else {
            gx_ += vt_x * P.dt; gy_ += vt_y * P.dt; gz_ += vt_z * P.dt;
          }
          S.tng_x[found_at] = gx_; S.tng_y[found_at] = gy_; S.tng_z[found_at] = gz_;
        }
        fx += fn * nx + ft_x; fy += fn * ny + ft_y; fz += fn * nz + ft_z;
        tx += (ny * ft_z - nz * ft_y) * a_i;
        ty += (nz * ft_x - nx * ft_z) * a_i;
        tz += (nx * ft_y - ny * ft_x) * a_i;
      }
    }
  S.total_tng[i] = tot;
  S.fx[i] = fx; S.fy[i] = fy; S.fz[i] = fz;
  S.torx[i] = tx; S.tory[i] = ty; S.torz[i] = tz;
  if (st && S.status) atomicOr(S.status, st);
}

__global__ void k_dem_stage(RbxDemScene S, int stage, double dt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.n_dest) return;
  if (stage == 2) {
    S.x[i] += dt * S.u[i];
    S.y[i] += dt * S.v[i];
    S.z[i] += dt * S.w[i];
  } else {
    const double dtb2 = 0.5 * dt;
    const double m_inverse = 1. / S.m[i];
    S.u[i] += dtb2 * S.fx[i] * m_inverse;
    S.v[i] += dtb2 * S.fy[i] * m_inverse;
    S.w[i] += dtb2 * S.fz[i] * m_inverse;
    const double I_inverse = 1. / S.moi[i];
    S.wx[i] += dtb2 * S.torx[i] * I_inverse;
    S.wy[i] += dtb2 * S.tory[i] * I_inverse;
    S.wz[i] += dtb2 * S.torz[i] * I_inverse;
  }
}

}  // namespace

extern "C" int rbx_contact_lvc(const RbxDemScene *scene, const RbxCells *cells,
                               const RbxParams *params, void *stream) {
  if (!scene || !cells || !params || scene->limit < 1) return RBX_ERR_INVALID;
  if (scene->n_dest <= 0) return RBX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int T = 128;
  k_dem_update<<<rbx_blocks(scene->n_dest, T), T, 0, st>>>(*scene);
  k_dem_force<<<rbx_blocks(scene->n_dest, T), T, 0, st>>>(*scene, *cells, *params);
  RBX_CHECK_LAUNCH();
  return RBX_OK;
}

extern "C" int rbx_dem_step(const RbxDemScene *scene, int stage, double dt, void *stream) {
  if (!scene || stage < 1 || stage > 3) return RBX_ERR_INVALID;
  if (scene->n_dest <= 0) return RBX_OK;
  k_dem_stage<<<rbx_blocks(scene->n_dest, 256), 256, 0, (cudaStream_t)stream>>>(*scene, stage, dt);
  RBX_CHECK_LAUNCH();
  return RBX_OK;
}
