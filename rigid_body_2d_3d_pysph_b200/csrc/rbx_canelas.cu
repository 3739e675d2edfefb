// Canelas (2016) SPH-DCDEM contact between the spheres (radius rad_s) that
// discretise rigid bodies and walls: Hertz normal force with Bui-2014 damping.
//
//   BodyForce.initialize                  rigid_body_common.py:115-125
//   RigidBodyCanelasRigidRigid.loop       rigid_body_common.py:244-442
//   RigidBodyCanelasRigidWall.loop        rigid_body_common.py:445-628
//
// No scheme of the reference wires these equations and their tangential part
// is commented out upstream; what remains is one pair loop.  One thread per
// rigid (destination) particle over the cell list of the source particles;
// a source that belongs to a rigid body takes the RigidRigid law (effective
// mass and radius of the pair), any other source the RigidWall law (mass and
// radius of the destination).  HBM-bound streaming plus a short gather per
// neighbour; kept simple, like the DEMScheme kernels.
#include "rbx_common.cuh"

namespace {

__global__ void k_canelas(RbxScene S, RbxCells C, RbxParams P, RbxCanelas K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.n_rigid) return;
  const RbxGridInfo gi = *C.info;
  const double px = S.x[i], py = S.y[i], pz = S.z[i], ph = S.h[i];
  const double ud = S.u[i], vd = S.v[i], wd = S.w[i];
  const double md = S.m[i], rd = K.rad_s[i];
  const int dem = S.dem_id[i];
  const double Md = S.total_mass[S.body[i]];
  const double tmp1 = (1. - K.nu[i] * K.nu[i]) / K.E[i];
  double fx = md * P.gx, fy = md * P.gy, fz = md * P.gz;   // BodyForce :122-125

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
        if (C.sdem[q] == dem) continue;                          // :262 / :463
        const double x0 = px - C.sx[q], x1 = py - C.sy[q], x2 = pz - C.sz[q];
        const double r2 = rbx_r2(x0, x1, x2);
        if (!(r2 < hi2 || r2 < rbx_h2(rs2, C.sh[q]))) continue;   // NNPS
        const double rij = sqrt(r2);
        const int g = C.gidx[q];
        const double rs = K.rad_s[g];
        double overlap = -1.;
        if (rij > 0) overlap = rd + rs - rij;                    // :264-265
        if (!(overlap > 0)) continue;
        const double rinv = 1.0 / rij;
        const double nx = x0 * rinv, ny = x1 * rinv, nz = x2 * rinv;
        const double vn = (ud - S.u[g]) * nx + (vd - S.v[g]) * ny + (wd - S.w[g]) * nz;
        const double tmp2 = (1. - K.nu[g] * K.nu[g]) / K.E[g];
        double m_eff = Md, r_eff = rd;                           // RigidWall :486-487
        const int bs = K.src_body[g];
        if (bs >= 0) {                                           // RigidRigid :293-295
          const double Ms = S.total_mass[bs];
          m_eff = Md * Ms / (Md + Ms);
          r_eff = rd * rs / (rd + rs);
        }
        const double E_eff = 1. / (tmp1 + tmp2);
        const double sr = sqrt(r_eff);
        const double kn = 4. / 3. * E_eff * sr;                  // :302
        const double gamma_n = K.Cn * sqrt(6. * m_eff * E_eff * sr);   // :305
        const double mag = kn * (overlap * sqrt(overlap)) - gamma_n * vn;
        fx += mag * nx; fy += mag * ny; fz += mag * nz;          // :307-309, :418-420
      }
    }
  S.fx[i] = fx; S.fy[i] = fy; S.fz[i] = fz;
}

}  // namespace

extern "C" int rbx_contact_canelas(const RbxScene *scene, const RbxCells *cells,
                                   const RbxParams *params, const RbxCanelas *canelas,
                                   void *stream) {
  if (!scene || !cells || !params || !canelas) return RBX_ERR_INVALID;
  if (!canelas->rad_s || !canelas->E || !canelas->nu || !canelas->src_body)
    return RBX_ERR_INVALID;
  if (scene->n_rigid <= 0) return RBX_OK;
  const int T = 128;
  k_canelas<<<rbx_blocks(scene->n_rigid, T), T, 0, (cudaStream_t)stream>>>(
      *scene, *cells, *params, *canelas);
  RBX_CHECK_LAUNCH();
  return RBX_OK;
}
