// Fused Mofidi-style contact evaluation with sparse (particle, source body)
// slots.  Replaces the five equation groups wired at
// /root/reference/code/rigid_body_3d.py:641-698:
//   ComputeContactForceNormals                  rigid_body_common.py:631-723
//   ComputeContactForceDistanceAndClosestPoint  rigid_body_common.py:726-836
//   BodyForce.initialize                        rigid_body_common.py:115-125
//   ComputeContactForce.post_loop               rigid_body_common.py:839-1032
//
// Two launches over the same decomposition: one CTA per "chunk" (<= 128
// consecutive particles of one rigid body, thread t <-> particle p0 + t).
//
// k_neighbours (issue bound; shared-memory staged; only on list rebuilds)
//   1. reduces the chunk's bounding box,
//   2. streams the cell-list rows overlapping box +- reach (coalesced SoA
//      loads, 4 independent loads per thread per barrier), drops the body's
//      own particles and everything outside the box, and compacts the rest
//      into a shared-memory tile (deterministic ballot/prefix compaction),
//   3. every thread tests its particle against the tile (broadcast
//      shared-memory reads; FP32 with a conservative threshold, because the
//      list only has to be a superset) and appends the hits to its neighbour
//      list in HBM, laid out [entry][particle] so that both this write and
//      the later read coalesce.
// k_slots (latency bound; no shared memory, no block barriers)
//   4. per distinct source body (ascending dem_id) the thread accumulates
//      the slot sums in registers, four list entries in flight at a time --
//      single pass: the distance sum of pass 2 is n . sum(XIJ m/rho W), so
//      the reference's two pair loops collapse into one -- then applies the
//      spring/dashpot/Coulomb law with the history carried in the sparse
//      slot table,
//   5. writes fx, fy, fz; k_bodies (rbx_bodies.cu) sums them per body with a
//      fixed shuffle tree (deterministic).
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

#ifndef RBX_KLD
#define RBX_KLD 4
#endif
#ifndef RBX_KACC
#define RBX_KACC 4
#endif
#ifndef RBX_NB_MINB
#define RBX_NB_MINB 6
#endif
#ifndef RBX_SLOTS_MINB
#define RBX_SLOTS_MINB 4
#endif
constexpr int kWarps = RBX_CHUNK / 32;
constexpr int kBatch = 4;        // staged candidates per thread per iteration
constexpr int kLd = RBX_KLD;     // list entries in flight per thread in k_slots

__global__ void __launch_bounds__(RBX_CHUNK, RBX_NB_MINB)
k_neighbours(RbxScene S, RbxCells C, RbxParams P, double reach) {
  // `reach` here is the LIST radius = neighbour reach + skin.  The list is a
  // superset of the neighbour set; k_slots applies the exact predicate.
  if (S.rebuild && *S.rebuild == 0u) return;      // lists still valid
  // tile entry: position relative to the box centre in FP32 (x, y, z) and the
  // global index of the source (w, as bits); dem_id beside it
  __shared__ float4 t_f[RBX_TILE];
  __shared__ int t_dem[RBX_TILE];
  __shared__ double red[kWarps][6];
  __shared__ int wtot[2][kBatch][kWarps];
  __shared__ int row_s[RBX_CHUNK], row_off[RBX_CHUNK + 1];
  __shared__ int wscan[kWarps];
  __shared__ int range[6];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const RbxGridInfo gi = *C.info;
  const size_t n_rigid = (size_t)S.n_rigid;
  // persistent CTAs over the chunks (a skipped evaluation then costs one
  // wave of CTAs, not one CTA per chunk), chunks handed out dynamically
  // through counters[7] (reset by k_list_clear) so that CTAs with dense
  // neighbourhoods do not hold up the tail
  __shared__ int s_chunk;
  for (;;) {
  __syncthreads();              // shared memory of the previous chunk is free
  if (tid == 0)
    s_chunk = S.counters ? (int)atomicAdd(&S.counters[7], 1ull) : -1;
  __syncthreads();
  const int chunk = s_chunk;
  if (chunk < 0 || chunk >= S.n_chunks) break;
  const int p0 = S.chunk_start[chunk], p1 = S.chunk_start[chunk + 1];
  const int p = p0 + tid;
  const bool valid = p < p1;
  const int my_dem = S.dem_id[p0];

  double px = 0, py = 0, pz = 0;
  if (valid) { px = S.x[p]; py = S.y[p]; pz = S.z[p]; }

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

  int nlist = 0;
  bool list_overflow = false;
  unsigned long long ncand = 0;
  const int cap = S.list_cap;

  // The list only has to be a SUPERSET of the neighbour set (k_slots applies
  // the exact FP64 predicate to every entry), so the candidate test runs in
  // FP32 -- twice the issue rate of FP64 on B200 and a 16-byte tile entry.
  // Coordinates are taken relative to the box centre: their FP32 rounding
  // error is <= 2^-24 E per component (E = half extent of the padded box), the
  // squared distance is off by < 16 * 2^-24 * (E + reach) * reach, which the
  // threshold absorbs.
  const double ccx = 0.5 * (blo[0] + bhi[0]), ccy = 0.5 * (blo[1] + bhi[1]),
               ccz = 0.5 * (blo[2] + bhi[2]);
  const double Eext = fmax(fmax(bhi[0] - ccx, bhi[1] - ccy), bhi[2] - ccz);
  const float thr_f = (float)((reach * reach + 32. * 5.96e-8 * (Eext + reach) * reach) *
                              (1. + 1e-6));
  const float pfx = (float)(px - ccx), pfy = (float)(py - ccy), pfz = (float)(pz - ccz);

  int tile_cnt = 0;
  int it = 0;

  // ---- 3. list predicate against the staged tile ---------------------------
  auto phase_a = [&]() {
    __syncthreads();  // tile complete
    if (valid) {
      ncand += (unsigned long long)tile_cnt;
#pragma unroll 4
      for (int j = 0; j < tile_cnt; j++) {
        const float4 f = t_f[j];
        const float dxf = pfx - f.x, dyf = pfy - f.y, dzf = pfz - f.z;
        if (fmaf(dzf, dzf, fmaf(dyf, dyf, dxf * dxf)) < thr_f) {
          if (nlist < cap) {
            S.nbr_pos[(size_t)nlist * n_rigid + p] = __float_as_int(f.w);
            S.nbr_dem[(size_t)nlist * n_rigid + p] = t_dem[j];
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

  // ---- 2. stage the cell rows overlapping the box --------------------------
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
      double sx[kBatch], sy[kBatch], sz[kBatch];
      int sd[kBatch], sq[kBatch], sg[kBatch];
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
          sg[k] = C.gidx[sq[k]];
          sx[k] = C.sx[sq[k]]; sy[k] = C.sy[sq[k]]; sz[k] = C.sz[sq[k]];
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
          t_f[dst] = make_float4((float)(sx[k] - ccx), (float)(sy[k] - ccy),
                                 (float)(sz[k] - ccz), __int_as_float(sg[k]));
          t_dem[dst] = sd[k];
        }
      }
      tile_cnt = run;
      it++;
    }
  }
  phase_a();

  if (valid) {
    S.nbr_cnt[p] = nlist;
    if (list_overflow && S.status) atomicOr(S.status, RBX_STATUS_LIST_OVERFLOW);
  }
  // counters: candidate distance tests, list entries written
  if (S.counters) {
    unsigned long long g = (unsigned long long)nlist, c = ncand;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      g += __shfl_xor_sync(0xffffffffu, g, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) { atomicAdd(&S.counters[2], c); atomicAdd(&S.counters[3], g); }
  }
  }  // chunk loop
}

// After a rebuild: remember where every body was, then lower the flag.
__global__ void k_list_commit(RbxScene S) {
  if (!S.rebuild || *S.rebuild == 0u || !S.xcm_ref) return;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= S.n_bodies) return;
#pragma unroll
  for (int j = 0; j < 3; j++) S.xcm_ref[3 * b + j] = S.xcm[3 * b + j];
#pragma unroll
  for (int k = 0; k < 9; k++) S.R_ref[9 * b + k] = S.R[9 * b + k];
}

__global__ void k_list_clear(RbxScene S, double skin) {
  if (S.counters) S.counters[7] = 0ull;      // chunk dispenser of k_neighbours
  // skin == 0: no reuse, the flag stays up and every evaluation rebuilds
  if (S.rebuild && S.xcm_ref && skin > 0.) *S.rebuild = 0u;
}

// ---- lists as the pair kernel reads them ------------------------------------
// On a rebuild, per window of kSortW consecutive particles: a stable counting
// sort of the particles by descending list length (work item t <-> particle
// nbr_order[t]), and every list rewritten into column t of nbr_srt with its
// entries grouped by source body in ascending dem_id (stable inside a body),
// bit 31 marking the first entry of a body.  k_slots then
//   * runs warps whose 32 lists have (nearly) the same length -- in particle
//     order the lengths range from 0 (interior) to 60+ (corners) inside one
//     warp and 60 % of the lanes idle,
//   * accumulates the sums of one source body in registers and knows a slot
//     is complete when the next marked entry arrives: no key search, no
//     shared-memory read-modify-write per pair.
constexpr int kSortW = 1024;
constexpr int kSortBins = 256;
constexpr unsigned kRunBit = 0x80000000u;

__global__ void __launch_bounds__(kSortW, 1)
k_list_sort(RbxScene S) {
  if (S.rebuild && *S.rebuild == 0u) return;      // lists still valid
  __shared__ int cnt[kSortW / 32][kSortBins];
  __shared__ int start[kSortBins];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int base = blockIdx.x * kSortW;
  const int p = base + tid;
  const bool valid = p < S.n_rigid;
  const int len = valid ? S.nbr_cnt[p] : 0;
  // bin 0 = threads past the end (sorted last), bin len + 1 otherwise
  const int key = valid ? (len < kSortBins - 2 ? len : kSortBins - 2) + 1 : 0;
  for (int i = tid; i < (kSortW / 32) * kSortBins; i += kSortW) (&cnt[0][0])[i] = 0;
  __syncthreads();
  const unsigned grp = __match_any_sync(0xffffffffu, key);
  const int within = __popc(grp & ((1u << lane) - 1u));
  if (within == 0) cnt[wid][key] = __popc(grp);
  __syncthreads();
  if (tid < kSortBins) {                // per bin: exclusive prefix over warps
    int run = 0;
#pragma unroll 4
    for (int w2 = 0; w2 < kSortW / 32; w2++) {
      const int c = cnt[w2][tid];
      cnt[w2][tid] = run;
      run += c;
    }
    start[tid] = run;
  }
  __syncthreads();
  if (wid == 0) {                       // descending exclusive scan over bins
    constexpr int per = kSortBins / 32;
    int tot[per], s = 0;
#pragma unroll
    for (int j = 0; j < per; j++) { tot[j] = start[kSortBins - 1 - (per * lane + j)]; s += tot[j]; }
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    int run = inc - s;
#pragma unroll
    for (int j = 0; j < per; j++) { start[kSortBins - 1 - (per * lane + j)] = run; run += tot[j]; }
  }
  __syncthreads();
  if (!valid) return;
  const int t = base + start[key] + cnt[wid][key] + within;
  const size_t n = (size_t)S.n_rigid;
  S.nbr_order[t] = p;
  const int *rd = S.nbr_dem + p, *rp = S.nbr_pos + p;
  int *out = S.nbr_srt + t;
  int cur = 0, o = 0;
  bool have = false;
  while (o < len) {
    int nk = 0x7fffffff;
    bool found = false;
    const int *c = rd;
    for (int e = 0; e < len; e++, c += n) {
      const int d = *c;
      if ((!have || d > cur) && d <= nk) { nk = d; found = true; }
    }
    if (!found) break;
    bool first = true;
    c = rd;
    const int *cp = rp;
    for (int e = 0; e < len; e++, c += n, cp += n) {
      if (*c == nk) {
        const unsigned q = (unsigned)*cp;
        out[(size_t)o * n] = (int)(first ? (q | kRunBit) : q);
        first = false;
        o++;
      }
    }
    cur = nk;
    have = true;
  }
  S.nbr_cnt_srt[t] = o;
}

constexpr int kAcc = RBX_KACC;  // completed slots parked in shared memory
constexpr int kFields = 8;      // ax ay az w1 bx by bz (qmin, qfirst)
constexpr int kOvf = 28;        // further slots parked in local memory

// Per-particle state of the force law that outlives one batch of slots.
struct SlotOut {
  double cfx, cfy, cfz;   // contact force on the particle so far
  int nout;               // history entries written
  int ki;                 // diagnostic slots written
  unsigned st;            // status bits
  unsigned nactive;       // slots in contact
};

// The parked slots of one particle -> normals, distance, force law, history
// (ascending dem_id by construction of the lists).  Deliberately not inlined:
// it runs once per particle after the pair loop (and, for a particle touching
// more bodies than there are parking slots, in the middle of it), and its
// live ranges must not be added to those of the pair loop.
__device__ __noinline__ void
finalize_slots(const RbxScene *Sp, const RbxParams *Pp, const RbxDiag *Dp,
               const double (*acc)[kFields][RBX_CHUNK], const double (*ovf)[kFields],
               int nk, int p, int tid, SlotOut *out) {
  const RbxScene &S = *Sp;
  const RbxParams &P = *Pp;
  const RbxDiag &D = *Dp;
  const size_t n_rigid = (size_t)S.n_rigid;
  const int body = S.body[p];
  const double spacing0 = S.spacing0[body];
  double cfx = out->cfx, cfy = out->cfy, cfz = out->cfz;
  int nout = out->nout, ki = out->ki;
  unsigned st = out->st, nactive = out->nactive;
  for (int sl = 0; sl < nk; sl++) {
    double a_ax, a_ay, a_az, a_w1, a_bx, a_by, a_bz;
    int2 pg;
    if (sl < kAcc) {
      pg = reinterpret_cast<const int2 *>(&acc[sl][7][tid])[0];
      a_ax = acc[sl][0][tid]; a_ay = acc[sl][1][tid]; a_az = acc[sl][2][tid];
      a_w1 = acc[sl][3][tid];
      a_bx = acc[sl][4][tid]; a_by = acc[sl][5][tid]; a_bz = acc[sl][6][tid];
    } else {
      const double *o = ovf[sl - kAcc];
      pg = reinterpret_cast<const int2 *>(&o[7])[0];
      a_ax = o[0]; a_ay = o[1]; a_az = o[2]; a_w1 = o[3];
      a_bx = o[4]; a_by = o[5]; a_bz = o[6];
    }
    if (pg.y < 0) continue;        // nothing of this body is in range now
    const int key = S.dem_id[pg.y];
    const double a_w2 = a_w1;
    // ComputeContactForceNormals.post_loop :705-723
    double nx = 0., ny = 0., nz = 0.;
    if (a_w1 > 1e-12) {
      nx = a_ax / a_w1; ny = a_ay / a_w1; nz = a_az / a_w1;
      const double magn = sqrt(nx * nx + ny * ny + nz * nz);
      nx /= magn; ny /= magn; nz /= magn;
    }
    // ...DistanceAndClosestPoint.post_loop :829-836, with
    // dist_tmp = sum (n.XIJ) tmp2 = n . sum XIJ tmp2
    double dist = 0.;
    if (a_w2 > 1e-12) dist = (nx * a_bx + ny * a_by + nz * a_bz) / a_w2;
    const int gmin = pg.x;                 // global index (-1: none)

    // ComputeContactForce.post_loop :906-1032
    double ovl_out = 0., ft0 = 0., ft1 = 0., ft2 = 0.;
    const double overlap = spacing0 - dist;
    if (overlap > 0. && overlap != spacing0) {
      double vxs = 0., vys = 0., vzs = 0.;
      if (gmin >= 0) { vxs = S.u[gmin]; vys = S.v[gmin]; vzs = S.w[gmin]; }
      // previous state of this slot
      double dl0 = 0., dl1 = 0., dl2 = 0., fn0 = 0., fn1 = 0., fn2 = 0.;
      for (int s2 = 0; s2 < S.ks; s2++) {
        const int hk = S.hist_key_in[(size_t)s2 * n_rigid + p];
        if (hk < 0) break;
        if (hk == key) {
          const size_t o = (size_t)s2 * n_rigid + p, pl = (size_t)S.ks * n_rigid;
          dl0 = S.hist_dlt_in[o]; dl1 = S.hist_dlt_in[pl + o]; dl2 = S.hist_dlt_in[2 * pl + o];
          fn0 = S.hist_fn_in[o]; fn1 = S.hist_fn_in[pl + o]; fn2 = S.hist_fn_in[2 * pl + o];
          break;
        }
      }
      const double md = S.m[p];
      const double ud = S.u[p], vd = S.v[p], wd = S.w[p];
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
      cfx += fn0 + ft0; cfy += fn1 + ft1; cfz += fn2 + ft2;      // :1030-1032
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
    // (else: the slot is zeroed, :1014-1027 -- an absent sparse slot)
    if (D.key) {
      if (ki < RBX_MAX_KEYS) {
        const size_t o = (size_t)ki * n_rigid + p;
        D.key[o] = key;
        if (D.closest) D.closest[o] = gmin;
        if (D.nx) { D.nx[o] = nx; D.ny[o] = ny; D.nz[o] = nz; }
        if (D.dist) D.dist[o] = dist;
        if (D.overlap) D.overlap[o] = ovl_out;
        if (D.ftx) { D.ftx[o] = ft0; D.fty[o] = ft1; D.ftz[o] = ft2; }
      } else {
        st |= RBX_STATUS_SLOT_OVERFLOW;
      }
    }
    ki++;
  }
  out->cfx = cfx; out->cfy = cfy; out->cfz = cfz;
  out->nout = nout; out->ki = ki; out->st = st; out->nactive = nactive;
}

template <int DIM, bool UNIFORM_H>
__global__ void __launch_bounds__(RBX_CHUNK, RBX_SLOTS_MINB)
k_slots(const __grid_constant__ RbxScene S, const __grid_constant__ RbxParams P,
        const __grid_constant__ RbxDiag D, double h_uniform) {
  // parked slots: [slot][field][thread] -> conflict-free.  Field 7 packs
  // (closest source, first source of the body) as two ints.
  __shared__ double acc[kAcc][kFields][RBX_CHUNK];

  // work item t <-> particle nbr_order[t] (k_list_sort): full warps of equal
  // list length.  The per-body force/torque sum is done by k_bodies.
  const int tid = threadIdx.x, lane = tid & 31;
  const int t = blockIdx.x * RBX_CHUNK + tid;
  const bool valid = t < S.n_rigid;
  const size_t n_rigid = (size_t)S.n_rigid;

  unsigned nactive = 0, npairs = 0;
  if (valid) {
    const int p = S.nbr_order[t];
    const int nlist = S.nbr_cnt_srt[t];
    const double px = S.x[p], py = S.y[p], pz = S.z[p];
    const double ph = S.h[p];
    const double hij_u = 0.5 * (ph + h_uniform);
    const double rmin0 = 4. * S.spacing0[S.body[p]];  // :765
    const double vol = S.m[p] / S.rho[p];
    const double rs2 = P.radius_scale * P.radius_scale;
    const double hi2 = rbx_h2(rs2, ph);
    const double hj2_u = rbx_h2(rs2, h_uniform);

    // the slot being accumulated (registers)
    double ax = 0., ay = 0., az = 0., w1 = 0.;     // sum XIJ*tmp1, sum tmp1*RIJ (:686-690)
    double bx = 0., by = 0., bz = 0.;              // sum XIJ*tmp2             (:807)
    double r2thr = rmin0 * rmin0;  // r2 of the closest source so far (:811-818)
    int qmin = -1;           // its global index
    int qfirst = -1;         // first list entry of this body (-1: no slot open)
    bool touched = false;    // some entry passed the neighbour predicate
    int nk = 0;              // slots parked in shared memory
    SlotOut so;
    so.cfx = so.cfy = so.cfz = 0.;
    so.nout = 0; so.ki = 0; so.st = 0u; so.nactive = 0u;

    // A finished slot is parked in shared memory; the ones past kAcc (a
    // particle near more than kAcc bodies: corners) in local memory.
    double ovf[kOvf][kFields];
    auto park = [&]() {
      if (nk < kAcc) {
        acc[nk][0][tid] = ax; acc[nk][1][tid] = ay; acc[nk][2][tid] = az;
        acc[nk][3][tid] = w1;
        acc[nk][4][tid] = bx; acc[nk][5][tid] = by; acc[nk][6][tid] = bz;
        reinterpret_cast<int2 *>(&acc[nk][7][tid])[0] = make_int2(qmin, touched ? qfirst : -1);
        nk++;
      } else if (nk < kAcc + kOvf) {
        double *o = ovf[nk - kAcc];
        o[0] = ax; o[1] = ay; o[2] = az; o[3] = w1; o[4] = bx; o[5] = by; o[6] = bz;
        reinterpret_cast<int2 *>(&o[7])[0] = make_int2(qmin, touched ? qfirst : -1);
        nk++;
      } else {
        so.st |= RBX_STATUS_SLOT_OVERFLOW;
      }
    };

    // software pipeline over the list: the list entry of e + 1 + kLd is being
    // loaded (coalesced stream from HBM) and the position of source e + 1 is
    // being gathered (L1/L2) while the pair math of entry e runs.
    int ql[kLd];
    const int *cl = S.nbr_srt + t;
#pragma unroll
    for (int j = 0; j < kLd; j++) {
      ql[j] = (1 + j < nlist) ? cl[(size_t)(1 + j) * n_rigid] : 0;
    }
    int qc = nlist > 0 ? cl[0] : 0;
    cl += (size_t)(1 + kLd) * n_rigid;
    double sx, sy, sz, sh = 0.;
    {
      const int qi = qc & 0x7fffffff;
      sx = S.x[qi]; sy = S.y[qi]; sz = S.z[qi];
      if (!UNIFORM_H) sh = S.h[qi];
    }

    // ---- pairs: every entry's pair math runs once and lands in the registers
    //      of the open slot; a marked entry parks the slot ---------------------
    {
      for (int e0 = 0; e0 < nlist; e0++) {
        // stage G for entry e0 + 1, stage L for entry e0 + 1 + kLd
        const int qn = ql[0];
        const int qni = qn & 0x7fffffff;
        const double gx = S.x[qni], gy = S.y[qni], gz = S.z[qni];
        double gh = 0.;
        if (!UNIFORM_H) gh = S.h[qni];
#pragma unroll
        for (int j = 0; j + 1 < kLd; j++) ql[j] = ql[j + 1];
        ql[kLd - 1] = (e0 + 1 + kLd < nlist) ? *cl : 0;
        cl += n_rigid;
        do {
          const int qi = qc & 0x7fffffff;
          if (qc < 0) {                        // first entry of a source body
            if (qfirst >= 0) park();
            ax = ay = az = w1 = bx = by = bz = 0.;
            r2thr = rmin0 * rmin0; qmin = -1; qfirst = qi; touched = false;
          }
          const double x0 = px - sx, x1 = py - sy, x2 = pz - sz;
          const double r2 = rbx_r2(x0, x1, x2);
          // exact neighbour predicate on the list entry:
          // the list was built with a skin, possibly several steps ago
          if (!(r2 < hi2 || r2 < (UNIFORM_H ? hj2_u : rbx_h2(rs2, sh)))) break;
          npairs++;
          touched = true;
          // 1/r from rsqrt (1 ulp) instead of sqrt + division: the sums
          // below move by a few ulp (tolerance 1e-10), the dependent FP64
          // chain per entry is 3x shorter.  The closest-point decision, which
          // must match the CPU path bit for bit, still compares correctly
          // rounded sqrt values (below).
          const double rinv = rsqrt(r2);
          const double rij = r2 * rinv;
          const double hij = UNIFORM_H ? hij_u : 0.5 * (ph + sh);
          const double wij = rbx_quintic<DIM>(rij, hij);
          const double tmp2 = vol * wij;                 // :803  m/rho * W
          const double tmp1 = tmp2 * rinv;               // :683  m/(rho r) * W
          ax += x0 * tmp1; ay += x1 * tmp1; az += x2 * tmp1;   // :686-688
          w1 += tmp2;                                    // :690  tmp1 * r
          bx += x0 * tmp2; by += x1 * tmp2; bz += x2 * tmp2;   // :807 (n . sum)
          // :809: the second weight sum equals the first (tmp1*r == tmp2)
          if (r2 <= r2thr * (1. + 1e-14)) {              // :811 (+ tie rule Q6)
            // possible new closest source: decide exactly as the reference
            // does, on correctly rounded distances
            const double rex = sqrt(r2);
            const double rmin = (qmin >= 0) ? sqrt(r2thr) : rmin0;
            bool take = rex < rmin;
            if (!take && qmin >= 0 && rex == rmin)     // exact tie: lowest
              take = qi < qmin;                        // global index wins
            if (take) { r2thr = r2; qmin = qi; }
          }
        } while (false);
        qc = qn; sx = gx; sy = gy; sz = gz; sh = gh;
      }
    }
    if (qfirst >= 0) park();
    finalize_slots(&S, &P, &D, acc, ovf, nk, p, tid, &so);
    nactive = so.nactive;
    if (so.nout < S.ks) S.hist_key_out[(size_t)so.nout * n_rigid + p] = -1;
    if (D.key)
      for (int k2 = so.ki; k2 < RBX_MAX_KEYS; k2++) D.key[(size_t)k2 * n_rigid + p] = -1;
    if (so.st && S.status) atomicOr(S.status, so.st);
    const double md = S.m[p];                        // BodyForce :122-125
    S.fx[p] = md * P.gx + so.cfx; S.fy[p] = md * P.gy + so.cfy; S.fz[p] = md * P.gz + so.cfz;
  }

  if (S.counters) {
    unsigned na = nactive, np_ = npairs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      na += __shfl_xor_sync(0xffffffffu, na, o);
      np_ += __shfl_xor_sync(0xffffffffu, np_, o);
    }
    if (lane == 0 && na) atomicAdd(&S.counters[1], (unsigned long long)na);
    if (lane == 0 && np_) atomicAdd(&S.counters[0], (unsigned long long)np_);
  }
}

}  // namespace

static int check_contact_args(const RbxScene *scene, const RbxCells *cells, const RbxParams *params) {
  if (!scene || !cells || !params) return RBX_ERR_INVALID;
  if (scene->ks < 1 || (scene->dim != 2 && scene->dim != 3)) return RBX_ERR_INVALID;
  if (!(params->reach > 0.) || scene->list_cap < 1) return RBX_ERR_INVALID;
  if (!scene->nbr_pos || !scene->nbr_dem || !scene->nbr_cnt || !scene->counters) return RBX_ERR_INVALID;
  if (!scene->nbr_srt || !scene->nbr_order || !scene->nbr_cnt_srt) return RBX_ERR_INVALID;
  return RBX_OK;
}

extern "C" int rbx_contact_neighbours(const RbxScene *scene, const RbxCells *cells,
                                      const RbxParams *params, void *stream_) {
  int rc = check_contact_args(scene, cells, params);
  if (rc) return rc;
  if (scene->n_chunks <= 0) return RBX_OK;
  cudaStream_t st = (cudaStream_t)stream_;
  const int nb = scene->n_chunks;
  if (params->skin < 0.) return RBX_ERR_INVALID;
  {
    const int grid = nb < 148 * RBX_NB_MINB * 4 ? nb : 148 * RBX_NB_MINB * 4;
    k_neighbours<<<grid, RBX_CHUNK, 0, st>>>(*scene, *cells, *params, params->reach + params->skin);
  }
  k_list_sort<<<rbx_blocks(scene->n_rigid, kSortW), kSortW, 0, st>>>(*scene);
  if (scene->rebuild)
    k_list_commit<<<rbx_blocks(scene->n_bodies, 256), 256, 0, st>>>(*scene);
  k_list_clear<<<1, 1, 0, st>>>(*scene, params->skin);
  RBX_CHECK_LAUNCH();
  return RBX_OK;
}

extern "C" int rbx_contact_slots(const RbxScene *scene, const RbxCells *cells,
                                 const RbxParams *params, const RbxDiag *diag, void *stream_) {
  int rc = check_contact_args(scene, cells, params);
  if (rc) return rc;
  if (scene->n_chunks <= 0) return RBX_OK;
  cudaStream_t st = (cudaStream_t)stream_;
  RbxDiag d;
  if (diag) d = *diag; else memset(&d, 0, sizeof(d));
  const bool uni = params->h_uniform > 0.;
  const int ng = rbx_blocks(scene->n_rigid, RBX_CHUNK);
  if (scene->dim == 3) {
    if (uni) k_slots<3, true><<<ng, RBX_CHUNK, 0, st>>>(*scene, *params, d, params->h_uniform);
    else k_slots<3, false><<<ng, RBX_CHUNK, 0, st>>>(*scene, *params, d, 0.);
  } else {
    if (uni) k_slots<2, true><<<ng, RBX_CHUNK, 0, st>>>(*scene, *params, d, params->h_uniform);
    else k_slots<2, false><<<ng, RBX_CHUNK, 0, st>>>(*scene, *params, d, 0.);
  }
  RBX_CHECK_LAUNCH();
  return RBX_OK;
}

extern "C" int rbx_contact_mofidi(const RbxScene *scene, const RbxCells *cells,
                                  const RbxParams *params, const RbxDiag *diag, void *stream_) {
  int rc = rbx_contact_neighbours(scene, cells, params, stream_);
  if (rc) return rc;
  return rbx_contact_slots(scene, cells, params, diag, stream_);
}
