// Fused Mofidi-style contact evaluation with sparse (particle, source body)
// slots.  Replaces the five equation groups wired at
// /root/reference/code/rigid_body_3d.py:641-698:
//   ComputeContactForceNormals                  rigid_body_common.py:631-723
//   ComputeContactForceDistanceAndClosestPoint  rigid_body_common.py:726-836
//   BodyForce.initialize                        rigid_body_common.py:115-125
//   ComputeContactForce.post_loop               rigid_body_common.py:839-1032
//
// Three kernels.
//
// k_neighbours (instruction-issue bound; only on list rebuilds): persistent
//   CTAs, one "chunk" at a time (<= 128 consecutive particles of one rigid
//   body, thread t <-> particle p0 + t)
//   1. reduces the chunk's bounding box,
//   2. streams the cell-list rows overlapping box +- reach (coalesced SoA
//      loads, 4 independent loads per thread per barrier), drops the body's
//      own particles and everything outside the box, and compacts the rest
//      into a shared-memory tile (deterministic ballot/prefix compaction),
//   3. groups the tile by source body (stable counting sort in shared
//      memory), tests every particle against it (FP32 with a conservative
//      threshold: the list only has to be a superset) and appends the hits
//      to the particle's neighbour list in HBM, laid out [entry][particle].
// k_list_sort (HBM bound; only on list rebuilds): work items ordered by list
//   length per window of 256 particles, lists transposed into that order.
// k_slots (FP64 issue / latency bound; every step): work item <-> particle
//   4. walks its list, two entries per iteration, with the exact neighbour
//      predicate; the sums of one source body live in registers -- single
//      pass: the distance sum of the reference's second pair loop is
//      n . sum(XIJ m/rho W), so the two loops collapse into one -- and at a
//      body's last entry a contact prefilter either drops the slot or parks
//      it in shared memory,
//   5. finalize_slots: normals, distance, spring/dashpot/Coulomb law with the
//      history carried in the sparse slot table, for the parked slots in
//      ascending dem_id,
//   6. writes fx, fy, fz; k_reduce (rbx_bodies.cu) sums them per body with a
//      fixed shuffle tree (deterministic).
#include "rbx_common.cuh"
#include <string.h>

namespace {

// tuning knobs (variants: csrc/build.py build(defines=[...], out=...))
#ifndef RBX_KLD
#define RBX_KLD 2
#endif
#ifndef RBX_KACC
#define RBX_KACC 4
#endif
#ifndef RBX_NB_MINB
#define RBX_NB_MINB 6
#endif
#ifndef RBX_SLOTS_CTA
#define RBX_SLOTS_CTA 32
#endif
#ifndef RBX_SLOTS_MINB
#define RBX_SLOTS_MINB (512 / RBX_SLOTS_CTA)
#endif
constexpr int kWarps = RBX_CHUNK / 32;
constexpr int kBatch = 4;        // staged candidates per thread per iteration
constexpr int kLd = RBX_KLD;     // list entries in flight per thread in k_slots

constexpr int kHash = RBX_CHUNK;                 // hash slots = group ids of a tile
constexpr int kGIter = RBX_TILE / RBX_CHUNK;     // tile entries per thread
constexpr int kEmptyKey = (int)0x80000000;       // never a dem_id
constexpr unsigned kRunBit = 0x80000000u;
constexpr int kSplitBit = 0x40000000;           // in nbr_cnt: bodies may be split

__global__ void __launch_bounds__(RBX_CHUNK, RBX_NB_MINB)
k_neighbours(RbxScene S, RbxCells C, RbxParams P, double reach) {
  // `reach` here is the LIST radius = neighbour reach + skin.  The list is a
  // superset of the neighbour set; k_slots applies the exact predicate.
  if (S.rebuild && *S.rebuild == 0u) return;      // lists still valid
  // tile entry: position relative to the box centre in FP32 (x, y, z) and the
  // global index of the source (w, as bits); dem_id beside it.  r_*: in
  // arrival (cell) order; t_*: the same entries grouped by dem_id.
  __shared__ float4 r_f[RBX_TILE];
  __shared__ int r_dem[RBX_TILE];
  __shared__ float4 t_f[RBX_TILE];
  __shared__ int t_dem[RBX_TILE];
  __shared__ int h_key[kHash];
  __shared__ unsigned t_first[RBX_TILE / 32];   // per 32 tile entries: bit j = entry starts a source body
  __shared__ int g_off[kWarps][kHash];
  __shared__ int g_scan[kWarps];
  __shared__ double red[kWarps][6];
  __shared__ int wtot[2][kBatch][kWarps];
  __shared__ int row_s[RBX_CHUNK], row_off[RBX_CHUNK + 1];
  __shared__ int wscan[kWarps];
  __shared__ int range[6];
  __shared__ int s_shared_group;   // a tile held more source bodies than hash slots

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const RbxGridInfo gi = *C.info;
  const size_t n_rigid = (size_t)S.n_rigid;
  // persistent CTAs over the chunks (a skipped evaluation then costs one
  // wave of CTAs, not one CTA per chunk), chunks handed out dynamically
  // through counters[7] (reset by k_list_clear) so that CTAs with dense
  // neighbourhoods do not hold up the tail
  __shared__ int s_chunk;
  for (;;) {
  __syncthreads();              // shared memory of the previous chunk is free
  if (tid == 0) {
    s_chunk = S.counters ? (int)atomicAdd(&S.counters[7], 1ull) : -1;
    s_shared_group = 0;
  }
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
  int *wp = S.nbr_pos + (valid ? p : 0);   // next list entry of this particle

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
  int ntiles = 0;

  // ---- 3. a full tile: group it by source body, test it, append the hits ----
  auto phase_a = [&]() {
    __syncthreads();  // raw tile complete
    if (tile_cnt > 0) ntiles++;
    // (a) Entries of one source body become contiguous (stable inside a
    // body), so that a particle's hits come out as one run per body and
    // k_slots can sum a body in registers.  Group id = slot of the dem_id in
    // a 128-entry hash table; a stable counting sort over (group, warp) with
    // every warp owning a contiguous quarter of the tile.  Which slot a
    // dem_id gets may vary from run to run; the order inside a body does
    // not, and k_slots finalizes bodies in ascending dem_id.
    h_key[tid] = kEmptyKey;
#pragma unroll
    for (int w2 = 0; w2 < kWarps; w2++) g_off[w2][tid] = 0;
    __syncthreads();
    const int L = ((tile_cnt + RBX_CHUNK - 1) / RBX_CHUNK) * 32;   // per warp
    const int jb = wid * L;
    int gk[kGIter];
#pragma unroll
    for (int k = 0; k < kGIter; k++) {
      gk[k] = kHash + lane;               // absent entry: matches nobody
      if (32 * k < L) {
        const int j = jb + 32 * k + lane;
        if (j < tile_cnt) {
          const int d = r_dem[j];
          unsigned h = ((unsigned)d * 2654435761u) >> 25;
          bool placed = false;
          for (int probe = 0; probe < kHash; probe++) {
            const int old = atomicCAS(&h_key[h], kEmptyKey, d);
            if (old == kEmptyKey || old == d) { placed = true; break; }
            h = (h + 1u) & (kHash - 1);
          }
          // More than kHash source bodies in one tile (bodies of 1-4
          // particles): this body shares the group of another one and the
          // two come out as interleaved runs.  Every run is still one body
          // (the run marks come from t_dem), so flagging the chunk as split
          // makes k_filter keep and k_slots park every partial run, and
          // finalize_slots adds them up by dem_id.
          if (!placed) s_shared_group = 1;
          gk[k] = (int)h;
        }
        const unsigned grp = __match_any_sync(0xffffffffu, gk[k]);
        if (gk[k] < kHash && (grp & lt_mask) == 0u) g_off[wid][gk[k]] += __popc(grp);
        __syncwarp();
      }
    }
    __syncthreads();
    {                                      // counts -> first output position
      int c[kWarps], tot = 0;
#pragma unroll
      for (int w2 = 0; w2 < kWarps; w2++) { c[w2] = g_off[w2][tid]; tot += c[w2]; }
      int inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      if (lane == 31) g_scan[wid] = inc;
      __syncthreads();
      int run = inc - tot;
#pragma unroll
      for (int w2 = 0; w2 < kWarps; w2++) if (w2 < wid) run += g_scan[w2];
#pragma unroll
      for (int w2 = 0; w2 < kWarps; w2++) { g_off[w2][tid] = run; run += c[w2]; }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kGIter; k++) {
      if (32 * k < L) {
        const int j = jb + 32 * k + lane;
        const bool in = gk[k] < kHash;
        const unsigned grp = __match_any_sync(0xffffffffu, gk[k]);
        const int within = __popc(grp & lt_mask);
        const int base = in ? g_off[wid][gk[k]] : 0;
        __syncwarp();
        if (in && within == 0) g_off[wid][gk[k]] = base + __popc(grp);
        __syncwarp();
        if (in) {
          t_f[base + within] = r_f[j];
          t_dem[base + within] = r_dem[j];
        }
      }
    }
    __syncthreads();  // grouped tile complete
#pragma unroll
    for (int k = 0; k < kGIter; k++) {
      const int j = k * RBX_CHUNK + tid;
      const bool first = j < tile_cnt && (j == 0 || t_dem[j] != t_dem[j - 1]);
      const unsigned word = __ballot_sync(0xffffffffu, first);
      if (lane == 0) t_first[k * kWarps + wid] = word;
    }
    __syncthreads();
    // (b) list predicate.  32 tile entries at a time: a branch-free pass
    // collects the hits in a bit mask, a second loop appends them -- the
    // append then runs once per hit of the busiest lane instead of once per
    // tested entry in which any lane hits (4 in 5).  An entry is the global
    // index of the source, bit 31 set when it starts a new source body.
    if (valid) {
      ncand += (unsigned long long)tile_cnt;
      // `pending`: the next hit starts a new source body (always true for the
      // first hit of a tile: a body continued from the previous tile is a
      // split body, which k_slots adds up)
      bool pending = true;
      for (int j0 = 0; j0 < tile_cnt; j0 += 32) {
        unsigned mask = 0u;
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const float4 f = t_f[j0 + j];       // RBX_TILE is a multiple of 32
          const float dxf = pfx - f.x, dyf = pfy - f.y, dzf = pfz - f.z;
          if (fmaf(dzf, dzf, fmaf(dyf, dyf, dxf * dxf)) < thr_f) mask |= 1u << j;
        }
        const int left = tile_cnt - j0;
        if (left < 32) mask &= (1u << left) - 1u;
        const unsigned bounds = t_first[j0 >> 5];   // body boundaries of this block
        unsigned seen = 0u;                          // positions up to the previous hit
        const float *tw = &t_f[j0].w;
        while (mask) {
          const int b = __ffs((int)mask) - 1;
          mask &= mask - 1u;
          const unsigned upto = (2u << b) - 1u;      // positions 0..b (b = 31: all)
          const bool first = pending || (bounds & upto & ~seen) != 0u;
          seen = upto;
          pending = false;
          if (nlist < cap) {
            const unsigned q = (unsigned)__float_as_int(tw[4 * b]);
            *wp = (int)(first ? (q | kRunBit) : q);
            wp += n_rigid;
            nlist++;
          } else {
            list_overflow = true;
          }
        }
        pending = pending || (bounds & ~seen) != 0u;  // a body started after the last hit
      }
    }
    __syncthreads();  // tiles may be overwritten
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
      int add = 0;
#pragma unroll
      for (int k = 0; k < kBatch; k++)
#pragma unroll
        for (int w2 = 0; w2 < kWarps; w2++) add += wtot[buf][k][w2];
      if (tile_cnt + add > RBX_TILE) phase_a();   // uniform: the tile is full
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
          const int dst = off + __popc(bal[k] & lt_mask);
          r_f[dst] = make_float4((float)(sx[k] - ccx), (float)(sy[k] - ccy),
                                 (float)(sz[k] - ccz), __int_as_float(sg[k]));
          r_dem[dst] = sd[k];
        }
      }
      tile_cnt = run;
      it++;
    }
  }
  phase_a();

  if (valid) {
    // bit 30: the chunk took more than one tile, a source body may own
    // several runs of the list
    S.nbr_cnt[p] = nlist | ((ntiles > 1 || s_shared_group) ? kSplitBit : 0);
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

// ... and where every static particle (walls, halo) was
__global__ void k_static_commit(RbxScene S) {
  if (!S.rebuild || *S.rebuild == 0u || !S.static_ref) return;
  const int n = S.n_total - S.n_rigid;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int q = S.n_rigid + k;
    S.static_ref[3 * (size_t)k] = S.x[q];
    S.static_ref[3 * (size_t)k + 1] = S.y[q];
    S.static_ref[3 * (size_t)k + 2] = S.z[q];
  }
}

__global__ void k_list_clear(RbxScene S, double skin) {
  if (S.counters) S.counters[7] = 0ull;      // chunk dispenser of k_neighbours
  // skin == 0: no reuse, the flag stays up and every evaluation rebuilds
  if (S.rebuild && S.xcm_ref && skin > 0.) *S.rebuild = 0u;
}

// ---- lists as the pair kernel reads them ------------------------------------
// On a rebuild, per window of kSortW consecutive particles: a stable counting
// sort of the particles by descending list length (work item t <-> particle
// nbr_order[t]) and a transposition of the lists into that order (column t of
// nbr_srt), through shared memory so that both the read and the write
// coalesce.  The run marker moves from the first to the LAST entry of a
// source body on the way.  k_slots then
//   * runs warps whose 32 lists have (nearly) the same length -- in particle
//     order the lengths range from 0 (interior) to 60+ (corners) inside one
//     warp and 60 % of the lanes idle,
//   * accumulates the sums of one source body in registers and parks the
//     slot when the marked entry has been added: no key search, no
//     shared-memory read-modify-write per pair.
#ifndef RBX_SORT_W
#define RBX_SORT_W 256
#endif
#ifndef RBX_SORT_ROWS
#define RBX_SORT_ROWS 32
#endif
constexpr int kSortW = RBX_SORT_W;       // particles per window = threads per CTA
constexpr int kSortBins = 256;
constexpr int kSortRows = RBX_SORT_ROWS; // list rows staged in shared memory at a time
static_assert(kSortW >= kSortBins && kSortW % 32 == 0 && kSortW <= 1024, "kSortW");

__global__ void __launch_bounds__(kSortW, 1024 / kSortW)
k_list_sort(RbxScene S) {
  if (S.rebuild && *S.rebuild == 0u) return;      // lists still valid
  __shared__ int stage[kSortRows + 1][kSortW];    // [row][particle of the window]
  __shared__ int cnt[kSortW / 32][kSortBins];
  __shared__ int start[kSortBins];
  __shared__ int perm[kSortW];                    // sorted slot -> particle of the window
  __shared__ int lens[kSortW];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int base = blockIdx.x * kSortW;
  const int p = base + tid;
  const bool valid = p < S.n_rigid;
  const int len_raw = valid ? S.nbr_cnt[p] : 0;
  const int len = len_raw & (kSplitBit - 1);
  const size_t n = (size_t)S.n_rigid;

  // ---- 1. stable counting sort of the window by descending list length ------
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
#pragma unroll
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
    int tot[per], sum = 0;
#pragma unroll
    for (int j = 0; j < per; j++) { tot[j] = start[kSortBins - 1 - (per * lane + j)]; sum += tot[j]; }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    int run = inc - sum;
#pragma unroll
    for (int j = 0; j < per; j++) { start[kSortBins - 1 - (per * lane + j)] = run; run += tot[j]; }
  }
  __syncthreads();
  const int slot = start[key] + cnt[wid][key] + within;   // in [0, kSortW)
  perm[slot] = tid;
  lens[tid] = len_raw;
  if (valid) S.nbr_order[base + slot] = p;
  int maxlen = len;                     // window maximum (strip loop bound)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
  __syncthreads();                      // perm, lens complete; cnt is free
  if (lane == 0) cnt[0][wid] = maxlen;
  __syncthreads();
  maxlen = 0;
#pragma unroll
  for (int w2 = 0; w2 < kSortW / 32; w2++) maxlen = max(maxlen, cnt[0][w2]);

  // ---- 2. transpose, kSortRows list rows at a time (+ 1 row of lookahead) ----
  const int src = perm[tid];            // the particle whose list this thread writes out
  const int len_out_raw = lens[src];
  const int len_out = len_out_raw & (kSplitBit - 1);
  const int t = base + tid;             // ... into column t
  const int *rp = S.nbr_pos + p;
  for (int o0 = 0; o0 < maxlen; o0 += kSortRows) {
    const int r_in = min(len - o0, kSortRows + 1);
    {
      const int *c = rp + (size_t)o0 * n;
      for (int r = 0; r < r_in; r++, c += n) stage[r][tid] = *c;
    }
    __syncthreads();
    if (t < S.n_rigid) {
      const int r1 = min(len_out - o0, kSortRows);
      int *out = S.nbr_srt + (size_t)o0 * n + t;
      int v = r1 > 0 ? stage[0][src] : 0;
      for (int r = 0; r < r1; r++, out += n) {
        // last entry of a source body <=> the next entry starts one
        const bool more = o0 + r + 1 < len_out;
        const int nv = more ? stage[r + 1][src] : -1;
        *out = (int)(((unsigned)v & 0x7fffffffu) | (nv < 0 ? kRunBit : 0u));
        v = nv;
      }
    }
    __syncthreads();
  }
  if (t < S.n_rigid) S.nbr_cnt_srt[t] = len_out_raw;
}

constexpr int kSlotsCta = RBX_SLOTS_CTA;   // threads per CTA of k_slots
constexpr int kAcc = RBX_KACC;  // completed slots parked in shared memory
constexpr int kFields = 8;      // ax ay az w1 bx by bz (qmin, qany)
constexpr int kOvf = 28;        // further slots parked in local memory

// Per-particle state of the force law that outlives one batch of slots.
struct SlotOut {
  double cfx, cfy, cfz;   // contact force on the particle so far
  int nout;               // history entries written
  int ki;                 // diagnostic slots written
  unsigned st;            // status bits
  unsigned nactive;       // slots in contact
};

// The parked slots of one particle -> normals, distance, force law, history,
// in ascending dem_id.  A source body normally owns one slot; where its
// entries were split over several runs of the list (a chunk whose candidates
// did not fit one tile) the partial slots are added up here.  Deliberately
// not inlined: it runs once per particle after the pair loop and its live
// ranges must not be added to those of the loop.
__device__ __noinline__ void
finalize_slots(const RbxScene *Sp, const RbxParams *Pp, const RbxDiag *Dp,
               double (*acc)[kFields][kSlotsCta], double (*ovf)[kFields],
               int nk, int p, int tid, bool split, SlotOut *out) {
  const RbxScene &S = *Sp;
  const RbxParams &P = *Pp;
  const RbxDiag &D = *Dp;
  const size_t n_rigid = (size_t)S.n_rigid;
  const int body = S.body[p];
  const double spacing0 = S.spacing0[body];
  double cfx = out->cfx, cfy = out->cfy, cfz = out->cfz;
  int nout = out->nout, ki = out->ki;
  unsigned st = out->st, nactive = out->nactive;
  auto ids = [&](int sl) -> int2 & {
    return sl < kAcc ? reinterpret_cast<int2 *>(&acc[sl][7][tid])[0]
                     : reinterpret_cast<int2 *>(&ovf[sl - kAcc][7])[0];
  };
  auto fld = [&](int sl, int f) -> double {
    return sl < kAcc ? acc[sl][f][tid] : ovf[sl - kAcc][f];
  };
  // key of every slot: dem_id of one of its in-range sources (-1: none).
  // Quick reject first: dist = (n . B) / w with n = A / |A|, so a slot can be
  // in contact (overlap = spacing0 - dist > 0) only if A.B < spacing0 |A| w;
  // tested on squares with a 1e-9 margin (the exact path rounds at 1e-15) it
  // spares all but the few slots near contact everything that follows.  A
  // slot with w <= 1e-12 has n = 0, dist = 0, overlap == spacing0: inactive
  // too.  Not for split bodies (partial sums) and not for diagnostics, which
  // want every slot's normal and distance.
  const bool prefilter = !split && !D.key;
  bool any = false;
  for (int sl = 0; sl < nk; sl++) {
    int2 &pg = ids(sl);
    int src = pg.y;
    if (src >= 0 && prefilter) {
      const double a0 = fld(sl, 0), a1 = fld(sl, 1), a2 = fld(sl, 2), w = fld(sl, 3);
      const double ab = a0 * fld(sl, 4) + a1 * fld(sl, 5) + a2 * fld(sl, 6);
      const double aa = a0 * a0 + a1 * a1 + a2 * a2;
      const double lim = spacing0 * w;
      if (!(w > 1e-12) || (ab > 0. && ab * ab > lim * lim * aa * (1. + 1e-9))) src = -1;
    }
    pg.y = src >= 0 ? S.dem_id[src] : -1;
    any = any || src >= 0;
  }
  if (!any) return;
  int prev = -1;
  for (;;) {
    int key = 0x7fffffff;
    for (int sl = 0; sl < nk; sl++) {
      const int k2 = ids(sl).y;
      if (k2 > prev && k2 < key) key = k2;
    }
    if (key == 0x7fffffff) break;
    prev = key;
    double a_ax = 0., a_ay = 0., a_az = 0., a_w1 = 0., a_bx = 0., a_by = 0., a_bz = 0.;
    int gmin = -1;                       // closest source (global index, -1: none)
    for (int sl = 0; sl < nk; sl++) {
      const int2 pg = ids(sl);
      if (pg.y != key) continue;
      a_ax += fld(sl, 0); a_ay += fld(sl, 1); a_az += fld(sl, 2); a_w1 += fld(sl, 3);
      a_bx += fld(sl, 4); a_by += fld(sl, 5); a_bz += fld(sl, 6);
      if (pg.x >= 0) {
        if (gmin < 0) {
          gmin = pg.x;
        } else {                         // two partial slots: the closer one
          const double px = S.x[p], py = S.y[p], pz = S.z[p];
          const double ra = sqrt(rbx_r2(px - S.x[gmin], py - S.y[gmin], pz - S.z[gmin]));
          const double rb = sqrt(rbx_r2(px - S.x[pg.x], py - S.y[pg.x], pz - S.z[pg.x]));
          if (rb < ra || (rb == ra && pg.x < gmin)) gmin = pg.x;
        }
      }
    }
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

    // ComputeContactForce.post_loop :906-1032
    double ovl_out = 0., ft0 = 0., ft1 = 0., ft2 = 0.;
    const double overlap = spacing0 - dist;
    if (overlap > 0. && overlap != spacing0) {
      // velocity of particle q: u, v, w, or for a rigid particle under
      // RBX_PARAM_BODY_VEL the stage-1 velocity formed from its body
      auto velocity = [&](int q, double &uq, double &vq, double &wq) {
        if ((P.flags & RBX_PARAM_BODY_VEL) && q < S.n_rigid) {
          const int bq = S.body[q];
          rbx_point_velocity(S.R_prev + 9 * bq, S.omega + 3 * bq, S.vcm + 3 * bq, S.dx0[q],
                             S.dy0[q], S.dz0[q], uq, vq, wq);
        } else {
          uq = S.u[q]; vq = S.v[q]; wq = S.w[q];
        }
      };
      double vxs = 0., vys = 0., vzs = 0.;
      if (gmin >= 0) velocity(gmin, vxs, vys, vzs);
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
      double ud, vd, wd;
      velocity(p, ud, vd, wd);
      const double vij_x = ud - vxs, vij_y = vd - vys, vij_z = wd - vzs;
      const double vn = vij_x * nx + vij_y * ny + vij_z * nz;
      ovl_out = overlap;
      const double tmp = P.kr * overlap;
      double eta = 0.;
      if (S.eta_mode == 1) {                                          // :925
        const long long row = S.eta_row[body];   // < 0: array without a table
        if (row >= 0) eta = S.eta[row + key];
      }
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

__device__ __noinline__ int
park_overflow(double (*ovf)[kFields], int nk, double ax, double ay, double az, double w1,
              double bx, double by, double bz, int2 ids, unsigned *st) {
  if (nk >= kAcc + kOvf) { *st |= RBX_STATUS_SLOT_OVERFLOW; return nk; }
  double *o = ovf[nk - kAcc];
  o[0] = ax; o[1] = ay; o[2] = az; o[3] = w1; o[4] = bx; o[5] = by; o[6] = bz;
  reinterpret_cast<int2 *>(&o[7])[0] = ids;
  return nk + 1;
}

// One work item = one particle and the runs of its list that have to be
// evaluated exactly.  clist == nullptr: every particle, every run (the
// reference semantics in one pass; diagnostics, pair dump).  Otherwise item i
// of the compact list written by k_filter: {work item t, bit mask of the runs
// that the FP32 pass could not exclude (bit 31 = run 31 and every later one)}.
template <int DIM, bool UNIFORM_H, bool COMPACT>
__global__ void __launch_bounds__(kSlotsCta, RBX_SLOTS_MINB)
k_slots(const __grid_constant__ RbxScene S, const __grid_constant__ RbxParams P,
        const __grid_constant__ RbxDiag D, double h_uniform, int dense_thr) {
  // parked slots: [slot][field][thread] -> conflict-free.  Field 7 packs
  // (closest source, some in-range source of the body) as two ints.
  __shared__ double acc[kAcc][kFields][kSlotsCta];

  // work item t <-> particle nbr_order[t] (k_list_sort): full warps of equal
  // list length.  The per-body force/torque sum is done by k_reduce.
  const int tid = threadIdx.x, lane = tid & 31;
  const size_t n_rigid = (size_t)S.n_rigid;
  unsigned nactive = 0, npairs = 0;

  // Persistent CTAs, every one resident from the start, striding over the
  // blocks of kSlotsCta work items (the grid size is odd: a stride that is a
  // multiple of the 8 warps of a sort window would give a CTA the same length
  // class every time).  A finished warp takes its next block at once instead
  // of waiting for a CTA launch.
  // dense_thr > 0 (two-precision evaluation): the FP32 pass left
  // counters[6] particles.  While they are few the compact list is walked;
  // once a good part of the scene is in contact (a settled pile) the list is
  // scattered all over the work items and the pass over EVERY particle in
  // work-item order -- coalesced list rows, warps of equal list length --
  // is the faster one.  Both launches are issued, one of them returns here;
  // the results are the same either way (a run the first pass excluded adds
  // exactly nothing).
  if (dense_thr > 0 && (((int)S.counters[6] > dense_thr) == COMPACT)) return;
  const int nwork = COMPACT ? (int)S.counters[6] : S.n_rigid;
  const int nitems = (nwork + kSlotsCta - 1) / kSlotsCta;
  // particle and list length of the work item after this one: loaded a whole
  // item ahead, so that its particle data and first list rows can be pulled
  // into L2 while this item finishes (a warp otherwise starts every item with
  // three dependent DRAM round trips and nothing to overlap them with)
  int p_next = -1, cnt_next = 0, t_next = 0;
  unsigned mask_next = 0xffffffffu;
  int first_next = 0, last_next = 0x7fffffff;   // COMPACT: entries to walk
  auto fetch_item = [&](int i) {
    p_next = -1;
    if (i < nwork) {
      t_next = i;
      if (COMPACT) {
        const int4 c = reinterpret_cast<const int4 *>(S.clist)[i];
        t_next = c.x; mask_next = (unsigned)c.y; first_next = c.z; last_next = c.w;
      }
      p_next = S.nbr_order[t_next];
      cnt_next = S.nbr_cnt_srt[t_next];
    }
  };
  fetch_item(blockIdx.x * kSlotsCta + tid);
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
  const int t = t_next;
  const bool valid = p_next >= 0;
  const int p = p_next;
  const int cnt_raw = cnt_next;
  const unsigned run_mask = mask_next;
  const int e_first = COMPACT ? (first_next & 0xfffff) : 0;
  const int run_first = COMPACT ? (first_next >> 20) : 0;
  const int e_last = last_next;
  if (item + gridDim.x < nitems) fetch_item((item + gridDim.x) * kSlotsCta + tid);
  else p_next = -1;
  const int tn = t_next;
  if (valid) {
    // COMPACT: only the entries from the first to the last run that the FP32
    // pass kept (what lies outside was excluded and adds exactly nothing)
    const int ebeg = e_first;
    const int nlist = min(cnt_raw & (kSplitBit - 1), COMPACT ? e_last + 1 : 0x7fffffff);
    // partial slots of a split body and the slots of a diagnostics run are
    // all parked; otherwise only those that pass the contact prefilter
    const bool park_all = (cnt_raw & kSplitBit) != 0 || D.key != nullptr;
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
    bool touched = false;    // some entry passed the neighbour predicate
    int nk = 0;              // slots parked in shared memory
    int run = run_first;     // ordinal of the source-body run being read
    SlotOut so;
    so.cfx = so.cfy = so.cfz = 0.;
    so.nout = 0; so.ki = 0; so.st = 0u; so.nactive = 0u;

    // A finished slot is parked in shared memory; the ones past kAcc (a
    // particle near more than kAcc bodies: corners) in local memory.
    double ovf[kOvf][kFields];
    auto park = [&](int2 ids) {
      if (nk < kAcc) {
        acc[nk][0][tid] = ax; acc[nk][1][tid] = ay; acc[nk][2][tid] = az;
        acc[nk][3][tid] = w1;
        acc[nk][4][tid] = bx; acc[nk][5][tid] = by; acc[nk][6][tid] = bz;
        reinterpret_cast<int2 *>(&acc[nk][7][tid])[0] = ids;
        nk++;
      } else {
        nk = park_overflow(ovf, nk, ax, ay, az, w1, bx, by, bz, ids, &so.st);
      }
    };

    // Two list entries per iteration.  Their pair math is independent
    // straight-line code (no branch: a pair out of range has W = 0 and adds
    // nothing -- the quintic spline's support is the neighbour radius -- so
    // the predicate only gates the closest-point search), which doubles the
    // instruction-level parallelism of what is otherwise one long dependent
    // FP64 chain per entry; the sums are then added in list order.
    // Software pipeline: entries e + 2 + j (j < kLd) are loaded (coalesced
    // stream from HBM), the positions of sources e + 2, e + 3 are being
    // gathered (L1/L2) while entries e, e + 1 are processed.
    static_assert(kLd >= 2 && kLd % 2 == 0, "kLd");
    int ql[kLd];
    const int *cl = S.nbr_srt + t + (size_t)ebeg * n_rigid;
#pragma unroll
    for (int j = 0; j < kLd; j++) ql[j] = (ebeg + 2 + j < nlist) ? cl[(size_t)(2 + j) * n_rigid] : 0;
    int qc0 = ebeg < nlist ? cl[0] : 0;
    int qc1 = ebeg + 1 < nlist ? cl[n_rigid] : 0;
    cl += (size_t)(2 + kLd) * n_rigid;
    double c0x, c0y, c0z, c0h = 0., c1x, c1y, c1z, c1h = 0.;
    {
      const int qi = qc0 & 0x7fffffff, qj = qc1 & 0x7fffffff;
      c0x = S.x[qi]; c0y = S.y[qi]; c0z = S.z[qi];
      if (!UNIFORM_H) c0h = S.h[qi];
      c1x = S.x[qj]; c1y = S.y[qj]; c1z = S.z[qj];
      if (!UNIFORM_H) c1h = S.h[qj];
    }
    const double hmax2_u = fmax(hi2, hj2_u);
    auto pair_math = [&](double sx, double sy, double sz, double sh, double &x0, double &x1,
                         double &x2, double &r2, double &tmp1, double &tmp2, bool &in) {
      x0 = px - sx; x1 = py - sy; x2 = pz - sz;
      r2 = rbx_r2(x0, x1, x2);
      // exact neighbour predicate on the list entry: the
      // list was built with a skin, possibly several steps ago
      in = UNIFORM_H ? (r2 < hmax2_u) : (r2 < hi2 || r2 < rbx_h2(rs2, sh));
      // 1/r from rsqrt (1 ulp) instead of sqrt + division: the sums move by a
      // few ulp (tolerance 1e-10), the dependent FP64 chain is 3x shorter
      const double rinv = rbx_rsqrt(r2);
      const double rij = r2 * rinv;
      const double hij = UNIFORM_H ? hij_u : 0.5 * (ph + sh);
      const double wij = rbx_quintic_nb<DIM>(rij, hij);
      tmp2 = in ? vol * wij : 0.;                        // :803  m/rho * W
      tmp1 = tmp2 * rinv;                                // :683  m/(rho r) * W
    };
    auto add_pair = [&](int qc, double x0, double x1, double x2, double r2, double tmp1,
                        double tmp2, bool in) {
      const int qi = qc & 0x7fffffff;
      ax += x0 * tmp1; ay += x1 * tmp1; az += x2 * tmp1;     // :686-688
      w1 += tmp2;                                      // :690  tmp1 * r
      bx += x0 * tmp2; by += x1 * tmp2; bz += x2 * tmp2;     // :807 (n . sum)
      // :809: the second weight sum equals the first (tmp1*r == tmp2)
      if (in) {
        npairs++;
        touched = true;
        if (D.pairs) {                 // parity mode: the pair set as seen here
          const unsigned long long k2 = atomicAdd(D.pair_count, 1ull);
          if ((long long)k2 < D.pair_cap) { D.pairs[2 * k2] = p; D.pairs[2 * k2 + 1] = qi; }
          else so.st |= RBX_STATUS_PAIR_OVERFLOW;
        }
        // :811 closest source (+ tie rule Q6).  The reference compares
        // correctly rounded distances; a squared distance smaller by more
        // than a few ulp decides the same way without the square roots,
        // and only a near tie takes the exact path.
        bool take = r2 < r2thr * (1. - 1e-14);
        if (!take && r2 <= r2thr * (1. + 1e-14)) {
          const double rex = sqrt(r2);
          const double rmin = (qmin >= 0) ? sqrt(r2thr) : rmin0;
          take = rex < rmin;
          if (!take && qmin >= 0 && rex == rmin)       // exact tie: lowest
            take = qi < qmin;                          // global index wins
        }
        if (take) { r2thr = r2; qmin = qi; }
      }
      if (qc < 0) {                    // last entry of this source body
        // Contact prefilter on the finished sums (see finalize_slots): all
        // but the few slots near contact end here, in registers.
        bool keep = touched;
        if (keep && !park_all) {
          const double ab = ax * bx + ay * by + az * bz;
          const double aa = ax * ax + ay * ay + az * az;
          const double lim = (0.25 * rmin0) * w1;      // spacing0 * w
          keep = w1 > 1e-12 && !(ab > 0. && ab * ab > lim * lim * aa * (1. + 1e-9));
        }
        if (keep || park_all) park(make_int2(qmin, touched ? qi : -1));
        ax = ay = az = w1 = bx = by = bz = 0.;
        r2thr = rmin0 * rmin0; qmin = -1; touched = false;
      }
    };
    // COMPACT: is the run with this ordinal one of those to evaluate?
    auto selected = [&](int r) -> bool {
      return !COMPACT || ((run_mask >> (r < 31 ? r : 31)) & 1u) != 0u;
    };
#pragma unroll 2
    for (int e0 = ebeg; e0 < nlist; e0 += 2) {
      // stage G for entries e0 + 2, e0 + 3; stage L for e0 + 2 + kLd, + 3 + kLd
      const int qn0 = ql[0], qn1 = ql[1];
      const int i0 = qn0 & 0x7fffffff, i1 = qn1 & 0x7fffffff;
      const double g0x = S.x[i0], g0y = S.y[i0], g0z = S.z[i0];
      const double g1x = S.x[i1], g1y = S.y[i1], g1z = S.z[i1];
      double g0h = 0., g1h = 0.;
      if (!UNIFORM_H) { g0h = S.h[i0]; g1h = S.h[i1]; }
#pragma unroll
      for (int j = 0; j + 2 < kLd; j++) ql[j] = ql[j + 2];
      ql[kLd - 2] = (e0 + 2 + kLd < nlist) ? cl[0] : 0;
      ql[kLd - 1] = (e0 + 3 + kLd < nlist) ? cl[n_rigid] : 0;
      cl += 2 * n_rigid;
      // runs the FP32 pass has excluded are only stepped over (an excluded
      // slot adds exactly nothing: rigid_body_common.py:1014-1027)
      const bool s0 = selected(run);
      const int run1 = run + (qc0 < 0 ? 1 : 0);
      const bool s1 = e0 + 1 < nlist && selected(run1);
      if (s0 || s1) {
        double a0, a1, a2, ar, at1, at2, b0, b1, b2, br, bt1, bt2;
        bool ain, bin;
        pair_math(c0x, c0y, c0z, c0h, a0, a1, a2, ar, at1, at2, ain);
        pair_math(c1x, c1y, c1z, c1h, b0, b1, b2, br, bt1, bt2, bin);
        if (s0) add_pair(qc0, a0, a1, a2, ar, at1, at2, ain);
        if (s1) add_pair(qc1, b0, b1, b2, br, bt1, bt2, bin);
      }
      run = run1 + ((e0 + 1 < nlist && qc1 < 0) ? 1 : 0);
      qc0 = qn0; c0x = g0x; c0y = g0y; c0z = g0z; c0h = g0h;
      qc1 = qn1; c1x = g1x; c1y = g1y; c1z = g1z; c1h = g1h;
    }
    if (p_next >= 0) {
      rbx_prefetch_l2(S.x + p_next); rbx_prefetch_l2(S.y + p_next); rbx_prefetch_l2(S.z + p_next);
      rbx_prefetch_l2(S.h + p_next); rbx_prefetch_l2(S.m + p_next); rbx_prefetch_l2(S.rho + p_next);
      rbx_prefetch_l2(S.body + p_next);
      const int fb = COMPACT ? (first_next & 0xfffff) : 0;
      const int *ln = S.nbr_srt + tn + (size_t)fb * n_rigid;
      const int nl = (cnt_next & (kSplitBit - 1)) - fb;
#pragma unroll
      for (int k = 0; k < 2 + kLd; k++)
        if (k < nl) rbx_prefetch_l2(ln + (size_t)k * n_rigid);
    }
    if (nk > 0) finalize_slots(&S, &P, &D, acc, ovf, nk, p, tid, (cnt_raw & kSplitBit) != 0, &so);
    nactive += so.nactive;
    if (so.nactive && S.alist_out) {
      // owns history rows and a force that is not m g: see RbxScene.alist_out
      S.alist_out[atomicAdd(S.acount_out, 1u)] = p;
      if (S.body_tag) S.body_tag[S.body[p]] = 1;
    }
    if (so.nout < S.ks) S.hist_key_out[(size_t)so.nout * n_rigid + p] = -1;
    if (D.key)
      for (int k2 = so.ki; k2 < RBX_MAX_KEYS; k2++) D.key[(size_t)k2 * n_rigid + p] = -1;
    if (so.st && S.status) atomicOr(S.status, so.st);
    const double md = S.m[p];                        // BodyForce :122-125
    S.fx[p] = md * P.gx + so.cfx; S.fy[p] = md * P.gy + so.cfy; S.fz[p] = md * P.gz + so.cfz;
  }
  }  // work-item loop

  if (S.counters) {
    unsigned na = nactive, np_ = npairs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      na += __shfl_xor_sync(0xffffffffu, na, o);
      np_ += __shfl_xor_sync(0xffffffffu, np_, o);
    }
    if (lane == 0 && na) atomicAdd(&S.counters[1], (unsigned long long)na);
    // (COMPACT: the FP32 pass has counted the pairs of every run)
    if (!COMPACT && dense_thr == 0 && lane == 0 && np_)
      atomicAdd(&S.counters[0], (unsigned long long)np_);
  }
}


// ---- FP32 first pass ---------------------------------------------------------
// At config 5 fewer than 1 % of the (particle, source body) slots are in
// contact, but deciding that takes the slot's complete sums.  k_filter forms
// the same sums as k_slots in FP32 (positions relative to RbxScene.origin,
// one 16-byte gather per entry, twice the issue rate and half the registers of
// FP64) together with a RUNNING BOUND of their error, and drops a slot only if
// the contact condition of rigid_body_common.py:906-907,
//     overlap = spacing0 - dist > 0,   dist = (A . B) / (|A| w)
// (A = sum XIJ m/(rho r) W, B = sum XIJ m/rho W, w = sum m/rho W), fails for
// EVERY value the exact sums can have inside the bound.  What cannot be
// excluded is evaluated by the unchanged FP64 code (k_slots<COMPACT>), so the
// results are those of the one-pass FP64 evaluation bit for bit
// (tests/test_gpu_fast.py); a dropped slot contributes exactly nothing, as
// the zeroed slot of the reference's else branch (:1014-1027).
//
// Error model (u = 2^-24, E = largest |coordinate - origin| of the particle or
// any list neighbour, L = list radius, all per component):
//   stored coordinate      |d| <= u E         ->  XIJ:  ex = 2.1 u E
//   r (FMA sum, MUFU.RSQ)  er = sqrt(3) ex + 4 u L;   q = r / h:  er / h
//   W = T w(q):            |dW| <= T max|w'| er / h + 9 u W  (mean value
//                          theorem), and with dq = er / h, t_i clamped at 0:
//                          |w'| <= 5 (t3+dq)^4 + 30 (t2+dq)^4 + 75 (t1+dq)^4
//                               <= 1.34 (5 t3^4 + 30 t2^4 + 75 t1^4) + 1.5e5 dq^4
//                          ((a+b)^4 <= 1.1^3 a^4 + 11^3 b^4) =: D; an entry
//                          beyond the support adds only the dq^4 floor
//   XIJ / r:               (ex + er) / r
// so with the extra sums  SD = sum (T / h) D,  wA = sum tmp1:
// (SD = sum (T / h) D)
//   dw = er SD + cu w,   dA = sqrt3 ((ex + er) wA + dw),
//   dB = sqrt3 (ex w + L dw),   cu = (list_cap + 16) u  (accumulation).
// No contact is certain when
//   A.B - [w dB + L w dA + dA dB] > spacing0 (|A| + dA) (w + dw)
// (|A| <= w, |B| <= L w; |A| + dA is taken as the root of |A|^2 + dA (2 w +
// dA), MUFU.SQRT within the 1e-4 allowance on spacing0); NaN (coincident
// points) fails the test and is kept.  Runs past ordinal 30 share mask bit 31.
#ifndef RBX_FILTER_MINB
#define RBX_FILTER_MINB 32
#endif

// Two list entries per iteration, their pair math in PACKED FP32 (FFMA2 /
// FMUL2 / FADD2, rbx_common.cuh): the kernel is bound by instruction issue and
// by the L1 lookups of its position gathers, and entries e, e + 1 of a list
// run through identical arithmetic, so one instruction per operation serves
// both.  A run may end at either entry of an iteration; both cases share ONE
// closing block per iteration (the closing test at 3-6 active lanes was 35 %
// of the executed instructions when every entry carried its own): a lane
// whose run ends at the first entry keeps the second entry's weights out of
// the sums (exact zeros), closes, and then opens the next run with that
// entry.  Every operation is IEEE round-to-nearest in the order of the error
// model above.  Lr and cu come as kernel arguments: formed in the kernel,
// ptxas re-derives them from the FP64 parameters inside the closing block.
template <int DIM, bool UNIFORM_H>
__global__ void __launch_bounds__(32, RBX_FILTER_MINB)
k_filter(const __grid_constant__ RbxScene S, const __grid_constant__ RbxParams P,
         float h_uniform, float Lr, float cu) {
  const int lane = threadIdx.x;
  const size_t n_rigid = (size_t)S.n_rigid;
  const float4 *__restrict__ pos = reinterpret_cast<const float4 *>(S.pos32);
  // list entry -> position: the byte offset q * 16 in 32 bits (the shift drops
  // the run marker in bit 31; n_total < 2^28, checked by the caller)
  const char *__restrict__ posb = reinterpret_cast<const char *>(S.pos32);
  auto gather = [&](int q) -> float4 {
    return *reinterpret_cast<const float4 *>(posb + ((unsigned)q << 4));
  };
  const int nitems = (S.n_rigid + 31) / 32;
  unsigned npairs = 0;
  const bool dense_out = !S.alist_out || (P.flags & RBX_PARAM_DENSE_OUT);
  constexpr float kU = 5.9604645e-8f;            // 2^-24
  constexpr float kSqrt3 = 1.7320509f;
  const float sigma = DIM == 2 ? (float)(0.31830988618379067154 * 7.0 / 478.0)
                               : (float)(0.31830988618379067154 / 120.0);

  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int t = item * 32 + lane;
    const bool valid = t < S.n_rigid;
    unsigned mask = 0u;
    int efirst = 0, elast = 0, rfirst = 0;
    if (valid) {
      const int p = S.nbr_order[t];
      const int cnt_raw = S.nbr_cnt_srt[t];
      const int nlist = cnt_raw & (kSplitBit - 1);
      const bool all = (cnt_raw & kSplitBit) != 0;   // split bodies: partial sums
      const float4 me = pos[p];
      float s0, vol;
      if (S.aux32) {
        const float2 a = reinterpret_cast<const float2 *>(S.aux32)[p];
        vol = a.x; s0 = a.y;
      } else {
        s0 = (float)S.spacing0[S.body[p]];
        vol = (float)(S.m[p] / S.rho[p]);
      }
      // error coefficients of this particle
      const float E = fmaxf(fmaxf(fabsf(me.x), fabsf(me.y)), fabsf(me.z)) + Lr;
      const float ex = 2.1f * kU * E;
      const float er = kSqrt3 * ex + 4.f * kU * Lr;
      const float eu = ex + er;
      const float s0f = s0 * 1.0001f;
      // uniform h: constants of the kernel
      rbx_f2_t H1 = 0ull, TH = 0ull, TT = 0ull, DF = 0ull;
      if (UNIFORM_H) {
        const float h1 = 1.f / (0.5f * (me.w + h_uniform));
        const float T = vol * sigma * (DIM == 2 ? h1 * h1 : h1 * h1 * h1);
        const float dq = er * h1;
        const float dfloor = 1.5e5f * (dq * dq) * (dq * dq);
        H1 = rbx_f2(h1, h1); TT = rbx_f2(T, T); TH = rbx_f2(T * h1, T * h1);
        DF = rbx_f2(dfloor, dfloor);
      }

      // sums of the open run
      float ax = 0.f, ay = 0.f, az = 0.f, w1 = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
      float wA = 0.f, SD = 0.f;
      int run = 0, estart = 0;

      const int *cl = S.nbr_srt + t;
      int qa = nlist > 0 ? cl[0] : 0;
      int qb = nlist > 1 ? cl[n_rigid] : 0;
      int la = nlist > 2 ? cl[2 * n_rigid] : 0;
      int lb = nlist > 3 ? cl[3 * n_rigid] : 0;
      cl += 4 * n_rigid;
      float4 sa = gather(qa), sb = gather(qb);
      for (int e0 = 0; e0 < nlist; e0 += 2) {
        const rbx_f2_t dx = rbx_f2(me.x - sa.x, me.x - sb.x);
        const rbx_f2_t dy = rbx_f2(me.y - sa.y, me.y - sb.y);
        const rbx_f2_t dz = rbx_f2(me.z - sa.z, me.z - sb.z);
        if (!UNIFORM_H) {
          float ha, hb;
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ha) : "f"(0.5f * (me.w + sa.w)));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(hb) : "f"(0.5f * (me.w + sb.w)));
          H1 = rbx_f2(ha, hb);
          const rbx_f2_t vs = rbx_f2(vol * sigma, vol * sigma);
          TT = rbx_f2_mul(vs, DIM == 2 ? rbx_f2_mul(H1, H1) : rbx_f2_mul(rbx_f2_mul(H1, H1), H1));
          TH = rbx_f2_mul(TT, H1);
          const rbx_f2_t dq = rbx_f2_mul(rbx_f2(er, er), H1);
          const rbx_f2_t dq2 = rbx_f2_mul(dq, dq);
          DF = rbx_f2_mul(rbx_f2_mul(rbx_f2(1.5e5f, 1.5e5f), dq2), dq2);
        }
        // software pipeline: positions of the next two entries, list entries
        // of the two after them
        const int qa_now = qa, qb_now = qb;
        sa = gather(la); sb = gather(lb);
        qa = la; qb = lb;
        la = (e0 + 4 < nlist) ? cl[0] : 0;
        lb = (e0 + 5 < nlist) ? cl[n_rigid] : 0;
        cl += 2 * n_rigid;

        const rbx_f2_t r2 = rbx_f2_fma(dz, dz, rbx_f2_fma(dy, dy, rbx_f2_mul(dx, dx)));
        float r2a, r2b, ia, ib;
        rbx_f2_get(r2, r2a, r2b);
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ia) : "f"(r2a));
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ib) : "f"(r2b));
        const rbx_f2_t rinv = rbx_f2(ia, ib);
        const rbx_f2_t q = rbx_f2_mul(rbx_f2_mul(r2, rinv), H1);
        const rbx_f2_t mone = rbx_f2(-1.f, -1.f);
        float lo, hi;
        rbx_f2_get(rbx_f2_fma(q, mone, rbx_f2(3.f, 3.f)), lo, hi);
        const float t3a = fmaxf(lo, 0.f), t3b = fmaxf(hi, 0.f);
        const rbx_f2_t t3 = rbx_f2(t3a, t3b);
        rbx_f2_get(rbx_f2_fma(q, mone, rbx_f2(2.f, 2.f)), lo, hi);
        const rbx_f2_t t2 = rbx_f2(fmaxf(lo, 0.f), fmaxf(hi, 0.f));
        rbx_f2_get(rbx_f2_fma(q, mone, rbx_f2(1.f, 1.f)), lo, hi);
        const rbx_f2_t t1 = rbx_f2(fmaxf(lo, 0.f), fmaxf(hi, 0.f));
        const rbx_f2_t a3 = rbx_f2_mul(t3, t3), a2 = rbx_f2_mul(t2, t2), a1 = rbx_f2_mul(t1, t1);
        const rbx_f2_t b3 = rbx_f2_mul(a3, a3), b2 = rbx_f2_mul(a2, a2), b1 = rbx_f2_mul(a1, a1);
        const rbx_f2_t wv = rbx_f2_fma(rbx_f2(15.f, 15.f), rbx_f2_mul(b1, t1),
                                       rbx_f2_fma(rbx_f2(-6.f, -6.f), rbx_f2_mul(b2, t2),
                                                  rbx_f2_mul(b3, t3)));
        const rbx_f2_t Dq = rbx_f2_fma(rbx_f2(100.5f, 100.5f), b1,
                                       rbx_f2_fma(rbx_f2(40.2f, 40.2f), b2,
                                                  rbx_f2_fma(rbx_f2(6.7f, 6.7f), b3, DF)));
        const rbx_f2_t tmp2 = rbx_f2_mul(TT, wv);
        const rbx_f2_t tmp1 = rbx_f2_mul(tmp2, rinv);
        const rbx_f2_t sd = rbx_f2_mul(TH, Dq);

        const bool bvalid = e0 + 1 < nlist;
        const bool ca = qa_now < 0;            // the open run ends at entry e0
        const bool cb = qb_now < 0;            // ... or at e0 + 1 (0 when absent)
        if (t3a > 0.f) npairs++;
        if (t3b > 0.f && bvalid) npairs++;
        // weights of the second entry: out of the packed sums when it belongs
        // to the next run (or does not exist: then ca holds, the last entry
        // of a list always closes a run)
        float w1a, w1b, w2a, w2b, sda, sdb, dxa, dxb, dya, dyb, dza, dzb;
        rbx_f2_get(tmp1, w1a, w1b);
        rbx_f2_get(tmp2, w2a, w2b);
        rbx_f2_get(sd, sda, sdb);
        rbx_f2_get(dx, dxa, dxb); rbx_f2_get(dy, dya, dyb); rbx_f2_get(dz, dza, dzb);
        ax = fmaf(dxa, w1a, ax); ay = fmaf(dya, w1a, ay); az = fmaf(dza, w1a, az);
        bx = fmaf(dxa, w2a, bx); by = fmaf(dya, w2a, by); bz = fmaf(dza, w2a, bz);
        w1 += w2a; wA += w1a; SD += sda;
        {
          const float u1 = ca ? 0.f : w1b, u2 = ca ? 0.f : w2b, us = ca ? 0.f : sdb;
          ax = fmaf(dxb, u1, ax); ay = fmaf(dyb, u1, ay); az = fmaf(dzb, u1, az);
          bx = fmaf(dxb, u2, bx); by = fmaf(dyb, u2, by); bz = fmaf(dzb, u2, bz);
          w1 += u2; wA += u1; SD += us;
        }

        bool pend = ca || cb;
        bool redo = ca && bvalid;              // the second entry opens the next run
        int ecur = ca ? e0 : e0 + 1;
#pragma unroll 1
        while (pend) {                         // last entry of a source body
          const float dw = fmaf(er, SD, cu * w1);
          const float dA = kSqrt3 * fmaf(eu, wA, dw);
          const float dB = kSqrt3 * fmaf(ex, w1, Lr * dw);
          const float ab = fmaf(az, bz, fmaf(ay, by, ax * bx));
          const float aa = fmaf(az, az, fmaf(ay, ay, ax * ax));
          const float eab = 1.01f * fmaf(dA, dB, w1 * fmaf(Lr, dA, dB));
          const float lhs = ab - eab;
          const float aahi = fmaf(dA, fmaf(2.f, w1, dA), aa);
          const float wh = w1 + dw;
          // (not on squares: for a slot of a few entries at the edge of the
          // support w is 1e-12 .. 1e-11, A.B is 1e-25 and its square
          // underflows FP32 -- such slots, 60 % of what this pass used to
          // keep, could not be dropped although they are 3 spacings away)
          float sq;
          asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(aahi));
          const float rhs = s0f * wh * sq;
          // w <= 1e-12: the slot has no normal, dist = 0, overlap == spacing0
          const bool drop = (wh < 0.99e-12f) || (lhs > 0.f && lhs > rhs);
          if (!drop || all) {
            if (mask == 0u) { efirst = estart; rfirst = run < 31 ? run : 31; }
            elast = ecur;
            mask |= 1u << (run < 31 ? run : 31);
          }
          run++;
          estart = ecur + 1;
          // the next run starts empty, or with the second entry of this
          // iteration
          ax = ay = az = w1 = bx = by = bz = 0.f;
          wA = SD = 0.f;
          if (redo) {
            ax = dxb * w1b; ay = dyb * w1b; az = dzb * w1b;
            bx = dxb * w2b; by = dyb * w2b; bz = dzb * w2b;
            w1 = w2b; wA = w1b; SD = sdb;
          }
          pend = redo && cb;
          redo = false;
          ecur = e0 + 1;
        }
      }
      if (mask == 0u && dense_out) {
        // nothing can be in contact: BodyForce alone (:122-125), no history
        // (sparse outputs: both are already in place, RbxScene.alist_out)
        const double md = S.m[p];
        S.fx[p] = md * P.gx; S.fy[p] = md * P.gy; S.fz[p] = md * P.gz;
        S.hist_key_out[p] = -1;
      }
    }
    // the rest goes to the exact pass: {work item, run mask}, appended in
    // lane order (which block of the list a warp gets does not matter)
    const unsigned bal = __ballot_sync(0xffffffffu, mask != 0u);
    if (bal) {
      int base = 0;
      if (lane == 0) base = (int)atomicAdd(&S.counters[6], (unsigned long long)__popc(bal));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (mask != 0u) {
        const int k = base + __popc(bal & ((1u << lane) - 1u));
        reinterpret_cast<int4 *>(S.clist)[k] =
            make_int4(t, (int)mask, efirst | (rfirst << 20), elast);
      }
    }
  }
  unsigned np_ = npairs;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) np_ += __shfl_xor_sync(0xffffffffu, np_, o);
  if (lane == 0 && np_) atomicAdd(&S.counters[0], (unsigned long long)np_);
}

// Sparse outputs (RbxScene.alist_out): before the pair kernels run, take back
// what the last two evaluations left behind -- the force of the particles
// that were in contact at the previous evaluation becomes m g again
// (BodyForce, rigid_body_common.py:122-125), their bodies lose the tag, and
// the history rows of the out buffer, last written two evaluations ago, read
// "empty" (the zeroed slot of :1014-1027).
__global__ void k_sparse_reset(RbxScene S, RbxParams P) {
  const unsigned n_out = *S.acount_out, n_prev = *S.acount_prev;
  const unsigned n = n_out > n_prev ? n_out : n_prev;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (i < n_out) S.hist_key_out[S.alist_out[i]] = -1;
    if (i < n_prev) {
      const int p = S.alist_prev[i];
      const double md = S.m[p];
      S.fx[p] = md * P.gx; S.fy[p] = md * P.gy; S.fz[p] = md * P.gz;
      if (S.body_tag) S.body_tag[S.body[p]] = 0;
    }
  }
}

// pos32 of the particles [first, first + n)
__global__ void k_pos32(RbxScene S, int first, int n, int only_on_rebuild) {
  if (only_on_rebuild && S.rebuild && *S.rebuild == 0u) return;
  float4 *pos = reinterpret_cast<float4 *>(S.pos32);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int q = first + k;
    pos[q] = make_float4((float)(S.x[q] - S.origin[0]), (float)(S.y[q] - S.origin[1]),
                         (float)(S.z[q] - S.origin[2]), (float)S.h[q]);
  }
}

}  // namespace

static int check_contact_args(const RbxScene *scene, const RbxCells *cells, const RbxParams *params) {
  if (!scene || !cells || !params) return RBX_ERR_INVALID;
  if (scene->ks < 1 || (scene->dim != 2 && scene->dim != 3)) return RBX_ERR_INVALID;
  if (!(params->reach > 0.) || scene->list_cap < 1) return RBX_ERR_INVALID;
  if (!scene->nbr_pos || !scene->nbr_cnt || !scene->counters) return RBX_ERR_INVALID;
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
    const int cap = rbx_sm_count() * RBX_NB_MINB * 4;
    const int grid = nb < cap ? nb : cap;
    k_neighbours<<<grid, RBX_CHUNK, 0, st>>>(*scene, *cells, *params, params->reach + params->skin);
  }
  // FP32 positions of the static particles (walls, halo): whoever moves
  // them raises the rebuild flag
  if (scene->pos32 && scene->n_total > scene->n_rigid) {
    const int n = scene->n_total - scene->n_rigid;
    const int nbp = rbx_blocks(n, 256), capp = rbx_sm_count() * 8;
    k_pos32<<<nbp < capp ? nbp : capp, 256, 0, st>>>(*scene, scene->n_rigid, n, 1);
  }
  k_list_sort<<<rbx_blocks(scene->n_rigid, kSortW), kSortW, 0, st>>>(*scene);
  if (scene->rebuild)
    k_list_commit<<<rbx_blocks(scene->n_bodies, 256), 256, 0, st>>>(*scene);
  if (scene->rebuild && scene->static_ref && scene->n_total > scene->n_rigid) {
    const int nbs = rbx_blocks(scene->n_total - scene->n_rigid, 256), caps = rbx_sm_count() * 8;
    k_static_commit<<<nbs < caps ? nbs : caps, 256, 0, st>>>(*scene);
  }
  k_list_clear<<<1, 1, 0, st>>>(*scene, params->skin);
  RBX_CHECK_LAUNCH();
  return RBX_OK;
}

extern "C" int rbx_pos32_refresh(const RbxScene *scene, int32_t first, int32_t n,
                                 void *stream_) {
  if (!scene || first < 0 || n < 0 || (long long)first + n > scene->n_total) return RBX_ERR_INVALID;
  if (!scene->pos32 || n == 0) return RBX_OK;
  int nb = rbx_blocks(n, 256);
  const int cap = rbx_sm_count() * 16;
  k_pos32<<<nb < cap ? nb : cap, 256, 0, (cudaStream_t)stream_>>>(*scene, first, n, 0);
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
  if (d.pairs && (!d.pair_count || d.pair_cap < 0)) return RBX_ERR_INVALID;
  const bool uni = params->h_uniform > 0.;
  const int sms = rbx_sm_count();
  // two-precision evaluation unless the caller asks for the one-pass FP64
  // evaluation, wants per-slot diagnostics, or has no FP32 positions
  const bool fast = scene->pos32 && scene->clist && !(params->flags & RBX_PARAM_EXACT) &&
                    !d.key && !d.pairs && scene->list_cap < (1 << 20) &&
                    scene->n_total < (1 << 28);
  // all CTAs resident at once (RBX_SLOTS_MINB per SM), odd count
  int ng = rbx_blocks(scene->n_rigid, kSlotsCta);
  const int resident = sms * RBX_SLOTS_MINB - 1;
  if (ng > resident) ng = resident;
  if (scene->alist_out) {
    if (!scene->alist_prev || !scene->acount_out || !scene->acount_prev) return RBX_ERR_INVALID;
    k_sparse_reset<<<sms * 2, 256, 0, st>>>(*scene, *params);
    cudaMemsetAsync(scene->acount_out, 0, sizeof(uint32_t), st);
    if ((params->flags & RBX_PARAM_DENSE_OUT) && scene->body_tag)
      cudaMemsetAsync(scene->body_tag, 0, sizeof(int32_t) * (size_t)scene->n_bodies, st);
  }
  if (fast) {
    cudaMemsetAsync(&scene->counters[6], 0, sizeof(unsigned long long), st);
    int nf = rbx_blocks(scene->n_rigid, 32);
    const int resf = sms * RBX_FILTER_MINB - 1;
    if (nf > resf) nf = resf;
    const float hu = (float)params->h_uniform;
    const float f_lr = (float)((params->reach + params->skin) * (1. + 1e-6));
    const float f_cu = (float)(scene->list_cap + 16) * 5.9604645e-8f;
    if (scene->dim == 3) {
      if (uni) k_filter<3, true><<<nf, 32, 0, st>>>(*scene, *params, hu, f_lr, f_cu);
      else k_filter<3, false><<<nf, 32, 0, st>>>(*scene, *params, 0.f, f_lr, f_cu);
    } else {
      if (uni) k_filter<2, true><<<nf, 32, 0, st>>>(*scene, *params, hu, f_lr, f_cu);
      else k_filter<2, false><<<nf, 32, 0, st>>>(*scene, *params, 0.f, f_lr, f_cu);
    }
    // exact pass: over the compact list, or over every particle when the
    // list has grown past 1/8 of them (see k_slots)
    const int thr = scene->n_rigid / 8 > 0 ? scene->n_rigid / 8 : 1;
    if (scene->dim == 3) {
      if (uni) {
        k_slots<3, true, true><<<ng, kSlotsCta, 0, st>>>(*scene, *params, d, params->h_uniform, thr);
        k_slots<3, true, false><<<ng, kSlotsCta, 0, st>>>(*scene, *params, d, params->h_uniform, thr);
      } else {
        k_slots<3, false, true><<<ng, kSlotsCta, 0, st>>>(*scene, *params, d, 0., thr);
        k_slots<3, false, false><<<ng, kSlotsCta, 0, st>>>(*scene, *params, d, 0., thr);
      }
    } else {
      if (uni) {
        k_slots<2, true, true><<<ng, kSlotsCta, 0, st>>>(*scene, *params, d, params->h_uniform, thr);
        k_slots<2, true, false><<<ng, kSlotsCta, 0, st>>>(*scene, *params, d, params->h_uniform, thr);
      } else {
        k_slots<2, false, true><<<ng, kSlotsCta, 0, st>>>(*scene, *params, d, 0., thr);
        k_slots<2, false, false><<<ng, kSlotsCta, 0, st>>>(*scene, *params, d, 0., thr);
      }
    }
  } else if (scene->dim == 3) {
    if (uni) k_slots<3, true, false><<<ng, kSlotsCta, 0, st>>>(*scene, *params, d, params->h_uniform, 0);
    else k_slots<3, false, false><<<ng, kSlotsCta, 0, st>>>(*scene, *params, d, 0., 0);
  } else {
    if (uni) k_slots<2, true, false><<<ng, kSlotsCta, 0, st>>>(*scene, *params, d, params->h_uniform, 0);
    else k_slots<2, false, false><<<ng, kSlotsCta, 0, st>>>(*scene, *params, d, 0., 0);
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
