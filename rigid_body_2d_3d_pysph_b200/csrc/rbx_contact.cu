// Fused Mofidi-style contact evaluation with sparse (particle, source body)
// slots.  Replaces, in one launch, the five equation groups wired at
// /root/reference/code/rigid_body_3d.py:641-698:
//   ComputeContactForceNormals                  rigid_body_common.py:631-723
//   ComputeContactForceDistanceAndClosestPoint  rigid_body_common.py:726-836
//   BodyForce.initialize                        rigid_body_common.py:115-125
//   ComputeContactForce.post_loop               rigid_body_common.py:839-1032
//   SumUpExternalForces.reduce (chunk partials) rigid_body_common.py:128-175
//
// Work decomposition: one CTA per "chunk" (<= 128 consecutive particles of
// one rigid body).  The CTA
//   1. reduces the chunk's bounding box,
//   2. streams the cell-list rows overlapping box +- reach (coalesced SoA
//      loads), drops the body's own particles and everything outside the
//      box, and compacts the rest into a shared-memory tile (deterministic
//      ballot/prefix compaction, no atomics),
//   3. phase A: every thread tests its particle against the tile (broadcast
//      shared-memory reads, exact FP64 predicate) and appends hits to a
//      private neighbour list,
//   4. phase B: per distinct source body (ascending dem_id) the thread
//      accumulates the slot sums in registers -- single pass: the distance
//      sum of pass 2 is n . sum(XIJ m/rho W), so the reference's two pair
//      loops collapse into one -- then applies the spring/dashpot/Coulomb
//      law with the history carried in the sparse slot table,
//   5. warp-shuffle + shared-memory reduction of force and torque about the
//      body's centre of mass -> one partial per chunk (fixed order).
#include "rbx_common.cuh"
#include <string.h>

namespace {

struct SlotAcc {
  double ax, ay, az, w1;  // sum XIJ*tmp1, sum tmp1*RIJ      (:686-690)
  double bx, by, bz, w2;  // sum XIJ*tmp2, sum tmp2          (:807-809)
  double rmin;            // closest_point_dist_to_source    (:811-818)
  int pmin;               // sorted position of the closest source (-1 none)
  int gmin;               // its global index (tie rule: lowest index)
};

constexpr int kWarps = RBX_CHUNK / 32;
constexpr int kBatch = 4;   // staged candidates per thread per iteration

template <int DIM, bool UNIFORM_H>
__global__ void __launch_bounds__(RBX_CHUNK, 4)
k_contact(RbxScene S, RbxCells C, RbxParams P, RbxDiag D, double reach, double h_uniform) {
  __shared__ double t_x[RBX_TILE], t_y[RBX_TILE], t_z[RBX_TILE];
  __shared__ double t_h[UNIFORM_H ? 1 : RBX_TILE];
  __shared__ int t_pos[RBX_TILE], t_dem[RBX_TILE];
  __shared__ double red[kWarps][6];
  __shared__ int wtot[2][kBatch][kWarps];
  __shared__ int row_s[RBX_CHUNK], row_off[RBX_CHUNK + 1];
  __shared__ int wscan[kWarps];
  __shared__ int range[6];
  __shared__ unsigned long long cnt_s[3];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int chunk = blockIdx.x;
  const int p0 = S.chunk_start[chunk], p1 = S.chunk_start[chunk + 1];
  const int p = p0 + tid;
  const bool valid = p < p1;
  const int body = S.chunk_body[chunk];
  const int my_dem = S.dem_id[p0];
  const RbxGridInfo gi = *C.info;
  const int n_rigid = S.n_rigid;

  double px = 0, py = 0, pz = 0, ph = 0;
  if (valid) { px = S.x[p]; py = S.y[p]; pz = S.z[p]; ph = S.h[p]; }

  // ---- 1. chunk bounding box --------------------------------------------
  {
    double lo[3] = {valid ? px : 1e300, valid ? py : 1e300, valid ? pz : 1e300};
    double hi[3] = {valid ? px : -1e300, valid ? py : -1e300, valid ? pz : -1e300};
#pragma unroll
    for (int a = 0; a < 3; a++) { lo[a] = rbx_warp_min(lo[a]); hi[a] = rbx_warp_max(hi[a]); }
    if (lane == 0) {
#pragma unroll
      for (int a = 0; a < 3; a++) { red[wid][a] = lo[a]; red[wid][3 + a] = hi[a]; }
    }
    if (tid < 3) cnt_s[tid] = 0ull;
    __syncthreads();
  }
  double blo[3], bhi[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    blo[a] = red[0][a]; bhi[a] = red[0][3 + a];
#pragma unroll
    for (int w2 = 1; w2 < kWarps; w2++) {
      blo[a] = fmin(blo[a], red[w2][a]);
      bhi[a] = fmax(bhi[a], red[w2][3 + a]);
    }
    blo[a] -= reach; bhi[a] += reach;
  }
  if (tid == 0) {
    const double eps = 1e-9 * gi.cell;
    range[0] = rbx_cell_coord(blo[0] - eps, gi.x0, gi.inv_cell, gi.nx);
    range[1] = rbx_cell_coord(bhi[0] + eps, gi.x0, gi.inv_cell, gi.nx);
    range[2] = rbx_cell_coord(blo[1] - eps, gi.y0, gi.inv_cell, gi.ny);
    range[3] = rbx_cell_coord(bhi[1] + eps, gi.y0, gi.inv_cell, gi.ny);
    range[4] = rbx_cell_coord(blo[2] - eps, gi.z0, gi.inv_cell, gi.nz);
    range[5] = rbx_cell_coord(bhi[2] + eps, gi.z0, gi.inv_cell, gi.nz);
  }
  __syncthreads();
  const int cx0 = range[0], cx1 = range[1], cy0 = range[2], cy1 = range[3];
  const int cz0 = range[4], cz1 = range[5];
  const int nry = cy1 - cy0 + 1;
  const int nrows = nry * (cz1 - cz0 + 1);

  // ---- per-thread neighbour list (local memory, L1 resident) -------------
  int l_pos[RBX_LISTCAP];
  int l_dem[RBX_LISTCAP];
  int nlist = 0;
  bool list_overflow = false;
  unsigned long long ncand = 0;

  const double rs2 = P.radius_scale * P.radius_scale;
  const double hi2 = rbx_h2(rs2, ph);
  const double hj2_u = rbx_h2(rs2, h_uniform);

  int tile_cnt = 0;
  int it = 0;

  auto phase_a = [&]() {
    __syncthreads();  // tile complete
    if (valid) {
      ncand += (unsigned long long)tile_cnt;
      for (int j = 0; j < tile_cnt; j++) {
        const double r2 = rbx_r2(px - t_x[j], py - t_y[j], pz - t_z[j]);
        bool hit = r2 < hi2;
        if (!UNIFORM_H) hit = hit || (r2 < rbx_h2(rs2, t_h[j]));
        else hit = hit || (r2 < hj2_u);
        if (hit) {
          if (nlist < RBX_LISTCAP) {
            l_pos[nlist] = t_pos[j];
            l_dem[nlist] = t_dem[j];
            nlist++;
          } else {
            list_overflow = true;
          }
        }
      }
    }
    __syncthreads();  // tile may be overwritten
    tile_cnt = 0;
  };

  // ---- 2./3. stage the cell rows overlapping the box ----------------------
  // Rows (cy, cz) are contiguous index ranges of the sorted arrays.  All row
  // bounds are fetched at once (one row per thread), prefix-summed into one
  // flat candidate sequence, and that sequence is streamed kBatch*128
  // entries per iteration so that every thread has kBatch independent loads
  // in flight per barrier.
  for (int rb = 0; rb < nrows; rb += RBX_CHUNK) {
    const int r = rb + tid;
    int s0 = 0, len = 0;
    if (r < nrows) {
      const int cy = cy0 + r % nry, cz = cz0 + r / nry;
      const int row = (cz * gi.ny + cy) * gi.nx;
      s0 = C.cell_start[row + cx0];
      len = C.cell_start[row + cx1 + 1] - s0;
    }
    int inc = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    __syncthreads();            // previous users of row_* / wscan are done
    if (lane == 31) wscan[wid] = inc;
    __syncthreads();
    int woff = 0;
#pragma unroll
    for (int w2 = 0; w2 < kWarps; w2++) if (w2 < wid) woff += wscan[w2];
    row_s[tid] = s0;
    row_off[tid] = woff + inc - len;
    if (tid == RBX_CHUNK - 1) row_off[RBX_CHUNK] = woff + inc;
    __syncthreads();
    const int total = row_off[RBX_CHUNK];

    for (int base = 0; base < total; base += kBatch * RBX_CHUNK) {
      if (tile_cnt + kBatch * RBX_CHUNK > RBX_TILE) phase_a();
      bool keep[kBatch];
      double sx[kBatch], sy[kBatch], sz[kBatch], sh[kBatch];
      int sd[kBatch], sq[kBatch];
#pragma unroll
      for (int k = 0; k < kBatch; k++) {
        const int f = base + k * RBX_CHUNK + tid;
        keep[k] = false;
        sq[k] = -1;
        if (f < total) {
          // largest r with row_off[r] <= f  (row_off is non-decreasing)
          int lo = 0, hi = RBX_CHUNK - 1;
#pragma unroll
          for (int stp = 0; stp < 7; stp++) {
            const int mid = (lo + hi + 1) >> 1;
            if (row_off[mid] <= f) lo = mid; else hi = mid - 1;
          }
          sq[k] = row_s[lo] + (f - row_off[lo]);
        }
      }
#pragma unroll
      for (int k = 0; k < kBatch; k++) {
        if (sq[k] >= 0) {
          sd[k] = C.sdem[sq[k]];
          sx[k] = C.sx[sq[k]]; sy[k] = C.sy[sq[k]]; sz[k] = C.sz[sq[k]];
          if (!UNIFORM_H) sh[k] = C.sh[sq[k]];
        }
      }
      const int buf = it & 1;
      unsigned bal[kBatch];
#pragma unroll
      for (int k = 0; k < kBatch; k++) {
        if (sq[k] >= 0)
          keep[k] = (sd[k] != my_dem) && sx[k] >= blo[0] && sx[k] <= bhi[0] &&
                    sy[k] >= blo[1] && sy[k] <= bhi[1] && sz[k] >= blo[2] && sz[k] <= bhi[2];
        bal[k] = __ballot_sync(0xffffffffu, keep[k]);
        if (lane == 0) wtot[buf][k][wid] = __popc(bal[k]);
      }
      __syncthreads();
      int run = tile_cnt;
#pragma unroll
      for (int k = 0; k < kBatch; k++) {
        int off = run;
#pragma unroll
        for (int w2 = 0; w2 < kWarps; w2++) {
          const int c = wtot[buf][k][w2];
          if (w2 < wid) off += c;
          run += c;
        }
        if (keep[k]) {
          const int dst = off + __popc(bal[k] & ((1u << lane) - 1u));
          t_x[dst] = sx[k]; t_y[dst] = sy[k]; t_z[dst] = sz[k];
          if (!UNIFORM_H) t_h[dst] = sh[k];
          t_pos[dst] = sq[k]; t_dem[dst] = sd[k];
        }
      }
      tile_cnt = run;
      it++;
    }
  }
  phase_a();

  // ---- 4. phase B: slots, ascending source dem_id --------------------------
  double fx = 0, fy = 0, fz = 0;
  unsigned nactive = 0;
  if (valid) {
    const double md = S.m[p], rhod = S.rho[p];
    const double ud = S.u[p], vd = S.v[p], wd = S.w[p];
    const double spacing0 = S.spacing0[body];
    fx = md * P.gx; fy = md * P.gy; fz = md * P.gz;  // BodyForce :122-125
    unsigned st = list_overflow ? RBX_STATUS_LIST_OVERFLOW : 0u;

    int key = 0x7fffffff;
    for (int e = 0; e < nlist; e++) key = min(key, l_dem[e]);
    int nout = 0, ki = 0;
    while (key != 0x7fffffff) {
      int next = 0x7fffffff;
      SlotAcc a;
      a.ax = a.ay = a.az = a.w1 = a.bx = a.by = a.bz = a.w2 = 0.;
      a.rmin = 4. * spacing0;  // :765
      a.pmin = -1; a.gmin = 0x7fffffff;
      for (int e = 0; e < nlist; e++) {
        const int d = l_dem[e];
        if (d != key) {
          if (d > key && d < next) next = d;
          continue;
        }
        const int q = l_pos[e];
        const double x0 = px - C.sx[q], x1 = py - C.sy[q], x2 = pz - C.sz[q];
        const double rij = sqrt(rbx_r2(x0, x1, x2));
        const double hij = 0.5 * (ph + (UNIFORM_H ? h_uniform : C.sh[q]));
        const double wij = rbx_quintic<DIM>(rij, hij);
        const double tmp1 = md / (rhod * rij) * wij;   // :683
        a.ax += x0 * tmp1; a.ay += x1 * tmp1; a.az += x2 * tmp1;
        a.w1 += tmp1 * rij;                            // :690
        const double tmp2 = md / (rhod) * wij;         // :803
        a.bx += x0 * tmp2; a.by += x1 * tmp2; a.bz += x2 * tmp2;
        a.w2 += tmp2;                                  // :809
        if (rij <= a.rmin) {                           // :811 (+ tie rule Q6)
          const int g = C.gidx[q];
          if (rij < a.rmin || (a.pmin >= 0 && g < a.gmin)) {
            a.rmin = rij; a.pmin = q; a.gmin = g;
          }
        }
      }
      // ComputeContactForceNormals.post_loop :705-723
      double nx = 0., ny = 0., nz = 0.;
      if (a.w1 > 1e-12) {
        nx = a.ax / a.w1; ny = a.ay / a.w1; nz = a.az / a.w1;
        const double magn = sqrt(nx * nx + ny * ny + nz * nz);
        nx /= magn; ny /= magn; nz /= magn;
      }
      // ...DistanceAndClosestPoint.post_loop :829-836, with
      // dist_tmp = sum (n.XIJ) tmp2 = n . sum XIJ tmp2
      double dist = 0.;
      if (a.w2 > 1e-12) dist = (nx * a.bx + ny * a.by + nz * a.bz) / a.w2;
      double vxs = 0., vys = 0., vzs = 0.;
      if (a.pmin >= 0) { vxs = S.u[a.gmin]; vys = S.v[a.gmin]; vzs = S.w[a.gmin]; }

      // previous state of this slot
      double dl0 = 0., dl1 = 0., dl2 = 0., fn0 = 0., fn1 = 0., fn2 = 0.;
      for (int s = 0; s < S.ks; s++) {
        const int hk = S.hist_key_in[(size_t)s * n_rigid + p];
        if (hk < 0) break;
        if (hk == key) {
          const size_t o = (size_t)s * n_rigid + p, pl = (size_t)S.ks * n_rigid;
          dl0 = S.hist_dlt_in[o]; dl1 = S.hist_dlt_in[pl + o]; dl2 = S.hist_dlt_in[2 * pl + o];
          fn0 = S.hist_fn_in[o]; fn1 = S.hist_fn_in[pl + o]; fn2 = S.hist_fn_in[2 * pl + o];
          break;
        }
      }

      // ComputeContactForce.post_loop :906-1032
      double ovl_out = 0., ft0 = 0., ft1 = 0., ft2 = 0.;
      const double overlap = spacing0 - dist;
      bool active = false;
      if (overlap > 0. && overlap != spacing0) {
        active = true;
        const double vij_x = ud - vxs, vij_y = vd - vys, vij_z = wd - vzs;
        const double vn = vij_x * nx + vij_y * ny + vij_z * nz;
        ovl_out = overlap;
        const double tmp = P.kr * overlap;
        double eta = 0.;
        if (S.eta_mode == 1) eta = S.eta[S.eta_row[body] + key];       // :925
        else if (S.eta_mode == 2) eta = S.eta[0];
        eta = eta * sqrt(md / 2. * P.kr);                               // :926
        const double fnx = (tmp - eta * vn) * nx;
        const double fny = (tmp - eta * vn) * ny;
        const double fnz = (tmp - eta * vn) * nz;
        const double vij_magn = sqrt(vij_x * vij_x + vij_y * vij_y + vij_z * vij_z);
        if (vij_magn < 1e-12) {
          dl0 = dl1 = dl2 = 0.;   // fn (fn0..2) keeps its previous value: Q3
        } else {
          const double tx = vij_x - nx * vn, ty = vij_y - ny * vn, tz = vij_z - nz * vn;
          const double ti_magn = sqrt(tx * tx + ty * ty + tz * tz);
          double ti_x = 0., ti_y = 0., ti_z = 0.;
          if (ti_magn > 1e-12) { ti_x = tx / ti_magn; ti_y = ty / ti_magn; ti_z = tz / ti_magn; }
          const double sx_ = dl0 + vij_x * P.dt, sy_ = dl1 + vij_y * P.dt, sz_ = dl2 + vij_z * P.dt;
          const double ddt = sx_ * ti_x + sy_ * ti_y + sz_ * ti_z;
          dl0 = ddt * ti_x; dl1 = ddt * ti_y; dl2 = ddt * ti_z;
          const double fsx = -P.kf * dl0, fsy = -P.kf * dl1, fsz = -P.kf * dl2;
          const double ft_magn = sqrt(fsx * fsx + fsy * fsy + fsz * fsz);
          const double fn_magn = sqrt(fnx * fnx + fny * fny + fnz * fnz);
          const double ca = P.fric_coeff * fn_magn;
          const double ft_star = (ft_magn < ca) ? ft_magn : ca;  // (b<a)?b:a, App. C-7
          ft0 = -ft_star * ti_x; ft1 = -ft_star * ti_y; ft2 = -ft_star * ti_z;
          const double mx = -ft0 / P.kf, my = -ft1 / P.kf, mz = -ft2 / P.kf;
          const double lt = sqrt(mx * mx + my * my + mz * mz);
          dl0 = mx / lt; dl1 = my / lt; dl2 = mz / lt;            // Q1, Q2 (0/0 = NaN)
          fn0 = fnx; fn1 = fny; fn2 = fnz;
        }
      } else {
        dl0 = dl1 = dl2 = 0.; fn0 = fn1 = fn2 = 0.;
      }
      fx += fn0 + ft0; fy += fn1 + ft1; fz += fn2 + ft2;         // :1030-1032

      if (active) {
        nactive++;
        if (nout < S.ks) {
          const size_t o = (size_t)nout * n_rigid + p, pl = (size_t)S.ks * n_rigid;
          S.hist_key_out[o] = key;
          S.hist_dlt_out[o] = dl0; S.hist_dlt_out[pl + o] = dl1; S.hist_dlt_out[2 * pl + o] = dl2;
          S.hist_fn_out[o] = fn0; S.hist_fn_out[pl + o] = fn1; S.hist_fn_out[2 * pl + o] = fn2;
          nout++;
        } else {
          st |= RBX_STATUS_HIST_OVERFLOW;
        }
      }
      if (D.key) {
        if (ki < RBX_MAX_KEYS) {
          const size_t o = (size_t)ki * n_rigid + p;
          D.key[o] = key;
          if (D.closest) D.closest[o] = a.pmin >= 0 ? a.gmin : -1;
          if (D.nx) { D.nx[o] = nx; D.ny[o] = ny; D.nz[o] = nz; }
          if (D.dist) D.dist[o] = dist;
          if (D.overlap) D.overlap[o] = ovl_out;
          if (D.ftx) { D.ftx[o] = ft0; D.fty[o] = ft1; D.ftz[o] = ft2; }
        } else {
          st |= RBX_STATUS_SLOT_OVERFLOW;
        }
      }
      ki++;
      key = next;
    }
    if (nout < S.ks) S.hist_key_out[(size_t)nout * n_rigid + p] = -1;
    if (D.key)
      for (int k2 = ki; k2 < RBX_MAX_KEYS; k2++) D.key[(size_t)k2 * n_rigid + p] = -1;
    if (st && S.status) atomicOr(S.status, st);
    S.fx[p] = fx; S.fy[p] = fy; S.fz[p] = fz;
  }

  // ---- 5. chunk partial of SumUpExternalForces :158-175 --------------------
  double v6[6] = {0, 0, 0, 0, 0, 0};
  if (valid) {
    const double dx = px - S.xcm[3 * body], dy = py - S.xcm[3 * body + 1],
                 dz = pz - S.xcm[3 * body + 2];
    v6[0] = fx; v6[1] = fy; v6[2] = fz;
    v6[3] = dy * fz - dz * fy;
    v6[4] = dz * fx - dx * fz;
    v6[5] = dx * fy - dy * fx;
  }
#pragma unroll
  for (int a = 0; a < 6; a++) v6[a] = rbx_warp_sum(v6[a]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < 6; a++) red[wid][a] = v6[a];
  }
  // counters: one atomic per warp into shared, one per CTA into global
  {
    unsigned long long g = (unsigned long long)nlist, c = ncand, na = nactive;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      g += __shfl_xor_sync(0xffffffffu, g, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
      na += __shfl_xor_sync(0xffffffffu, na, o);
    }
    if (lane == 0) {
      atomicAdd(&cnt_s[0], g); atomicAdd(&cnt_s[1], na); atomicAdd(&cnt_s[2], c);
    }
  }
  __syncthreads();
  if (tid < 6) {
    double s = red[0][tid];
#pragma unroll
    for (int w2 = 1; w2 < kWarps; w2++) s += red[w2][tid];
    S.chunk_ft[(size_t)chunk * 6 + tid] = s;
  }
  if (tid < 3 && S.counters) atomicAdd(&S.counters[tid], cnt_s[tid]);
}

}  // namespace

extern "C" int rbx_contact_mofidi(const RbxScene *scene, const RbxCells *cells,
                                  const RbxParams *params, const RbxDiag *diag, void *stream_) {
  if (!scene || !cells || !params) return RBX_ERR_INVALID;
  if (scene->n_chunks <= 0) return RBX_OK;
  if (scene->ks < 1 || (scene->dim != 2 && scene->dim != 3)) return RBX_ERR_INVALID;
  if (!(params->reach > 0.)) return RBX_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream_;
  RbxDiag d;
  if (diag) d = *diag; else memset(&d, 0, sizeof(d));
  const bool uni = params->h_uniform > 0.;
  const int nb = scene->n_chunks;
  if (scene->dim == 3) {
    if (uni) k_contact<3, true><<<nb, RBX_CHUNK, 0, st>>>(*scene, *cells, *params, d, params->reach, params->h_uniform);
    else k_contact<3, false><<<nb, RBX_CHUNK, 0, st>>>(*scene, *cells, *params, d, params->reach, 0.);
  } else {
    if (uni) k_contact<2, true><<<nb, RBX_CHUNK, 0, st>>>(*scene, *cells, *params, d, params->reach, params->h_uniform);
    else k_contact<2, false><<<nb, RBX_CHUNK, 0, st>>>(*scene, *cells, *params, d, params->reach, 0.);
  }
  RBX_CHECK_LAUNCH();
  return RBX_OK;
}
