// Counting-sort cell list (replaces [upstream] LinkedListNNPS.update; SURVEY
// App. C-1) and the parity-mode neighbour enumeration.
//
// HBM-bound integer/byte work: every pass is a coalesced stream over the
// binned points (24 B in + 8 B out for the histogram pass, 4+4 B for the
// scatter, 36 B gather + 36 B store for the sorted SoA), grid sizes are a
// function of the point count only so the whole build is CUDA-graph safe
// (the grid dimensions live in RbxGridInfo on the device).
#include "rbx_common.cuh"

namespace {

struct BoundsWS {
  unsigned long long mn[3];  // ordered keys, memset 0xFF
  unsigned long long mx[3];  // ordered keys, memset 0x00
  unsigned int done;         // block counter, memset 0
  unsigned int pad;
};

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

#define RBX_SKIP_IF_VALID(C) do { if ((C).cond && *(C).cond == 0u) return; } while (0)

__global__ void k_bounds(RbxPoints P, RbxCells C, double min_cell, BoundsWS *ws,
                         uint32_t *status, int32_t *counts, int nscan) {
  RBX_SKIP_IF_VALID(C);
  // the cell histogram of k_count starts from zero: cleared here (only when
  // the list is rebuilt) rather than by a memset of cap_cells words per step
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nscan; k += gridDim.x * blockDim.x)
    counts[k] = 0;
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < P.n; k += gridDim.x * blockDim.x) {
    int g = P.index ? P.index[k] : k;
    double v[3] = {P.x[g], P.y[g], P.z[g]};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      mn[a] = fmin(mn[a], v[a]);
      mx[a] = fmax(mx[a], v[a]);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
    mn[a] = rbx_warp_min(mn[a]);
    mx[a] = rbx_warp_max(mx[a]);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      atomicMin(&ws->mn[a], rbx_ord(mn[a]));
      atomicMax(&ws->mx[a], rbx_ord(mx[a]));
    }
  }
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicAdd(&ws->done, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    RbxGridInfo gi;
    double lo[3], hi[3];
    for (int a = 0; a < 3; a++) {
      lo[a] = rbx_unord(atomicAdd(&ws->mn[a], 0ull));
      hi[a] = rbx_unord(atomicAdd(&ws->mx[a], 0ull));
    }
    if (P.n == 0) { lo[0] = lo[1] = lo[2] = 0.; hi[0] = hi[1] = hi[2] = 0.; }
    double cell = min_cell;
    long long nx, ny, nz;
    bool coarse = false;
    for (;;) {
      nx = (long long)floor((hi[0] - lo[0]) / cell) + 1;
      ny = (long long)floor((hi[1] - lo[1]) / cell) + 1;
      nz = (long long)floor((hi[2] - lo[2]) / cell) + 1;
      if (nx * ny * nz <= (long long)C.cap_cells) break;
      cell *= 1.25;
      coarse = true;
    }
    if (coarse && status) atomicOr(status, RBX_STATUS_GRID_COARSENED);
    gi.x0 = lo[0]; gi.y0 = lo[1]; gi.z0 = lo[2];
    gi.cell = cell; gi.inv_cell = 1.0 / cell;
    gi.nx = (int)nx; gi.ny = (int)ny; gi.nz = (int)nz;
    gi.ncells = (int)(nx * ny * nz);
    gi.npoints = P.n;
    gi.pad_ = 0;
    *C.info = gi;
  }
}

// The per-point passes use capped grids with grid-stride loops: when the
// neighbour lists are still valid (cond == 0) a skipped build costs a few
// hundred CTAs per kernel instead of one per 256 points.
__global__ void k_count(RbxPoints P, RbxCells C, int32_t *counts) {
  RBX_SKIP_IF_VALID(C);
  const RbxGridInfo gi = *C.info;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < P.n; k += gridDim.x * blockDim.x) {
    int g = P.index ? P.index[k] : k;
    int cx = rbx_cell_coord(P.x[g], gi.x0, gi.inv_cell, gi.nx);
    int cy = rbx_cell_coord(P.y[g], gi.y0, gi.inv_cell, gi.ny);
    int cz = rbx_cell_coord(P.z[g], gi.z0, gi.inv_cell, gi.nz);
    int c = (cz * gi.ny + cy) * gi.nx + cx;
    C.cell_of[k] = c;
    C.rank[k] = atomicAdd(&counts[c], 1);
  }
}

// ---- exclusive scan of n int32 (three passes, fixed launch geometry) ----
__global__ void k_scan_tiles(RbxCells C, const int32_t *in, int32_t *out, int32_t *tile_sum, int n) {
  RBX_SKIP_IF_VALID(C);
  __shared__ int32_t wsum[kScanThreads / 32];
  int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int32_t v[kScanItems];
  int32_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  // inclusive warp scan of the per-thread sums
  int32_t inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if ((threadIdx.x & 31) >= o) inc += t;
  }
  if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
  __syncthreads();
  int32_t woff = 0;
  for (int wdx = 0; wdx < (threadIdx.x >> 5); wdx++) woff += wsum[wdx];
  int32_t run = woff + inc - s;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
  if (threadIdx.x == kScanThreads - 1) tile_sum[blockIdx.x] = run;
}

__global__ void k_scan_sums(RbxCells C, int32_t *tile_sum, int ntiles) {
  RBX_SKIP_IF_VALID(C);
  // one block; sequential carry over chunks of blockDim.x tiles
  __shared__ int32_t wsum[32];
  __shared__ int32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < ntiles; base += blockDim.x) {
    int i = base + threadIdx.x;
    int32_t v = (i < ntiles) ? tile_sum[i] : 0;
    int32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
    __syncthreads();
    int32_t woff = 0;
    for (int wdx = 0; wdx < (threadIdx.x >> 5); wdx++) woff += wsum[wdx];
    int32_t carry = carry_s;
    if (i < ntiles) tile_sum[i] = carry + woff + inc - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = carry + woff + inc;
    __syncthreads();
  }
}

__global__ void k_scan_add(RbxCells C, int32_t *out, const int32_t *tile_sum, int n) {
  RBX_SKIP_IF_VALID(C);
  int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int32_t off = tile_sum[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; i++)
    if (base + i < n) out[base + i] += off;
}

__global__ void k_scatter(RbxPoints P, RbxCells C) {
  RBX_SKIP_IF_VALID(C);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < P.n; k += gridDim.x * blockDim.x) {
    int g = P.index ? P.index[k] : k;
    C.gidx[C.cell_start[C.cell_of[k]] + C.rank[k]] = g;
  }
}

// Arrival order inside a cell depends on atomic timing; sorting each cell's
// few entries by global index makes the list (and every sum over it)
// reproducible, and gives the lowest-index tie rule for free.
__global__ void k_sort_cells(RbxCells C) {
  RBX_SKIP_IF_VALID(C);
  const int ncells = C.info->ncells;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += gridDim.x * blockDim.x) {
  int s = C.cell_start[c], e = C.cell_start[c + 1];
  int n = e - s;
  if (n < 2) continue;
  int32_t *a = C.gidx + s;
  if (n <= 64) {
    for (int i = 1; i < n; i++) {
      int32_t v = a[i];
      int j = i - 1;
      while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; j--; }
      a[j + 1] = v;
    }
  } else {  // heapsort: any size, in place
    for (int start = n / 2 - 1; start >= 0; start--) {
      int root = start;
      for (;;) {
        int child = 2 * root + 1;
        if (child >= n) break;
        if (child + 1 < n && a[child] < a[child + 1]) child++;
        if (a[root] >= a[child]) break;
        int32_t t = a[root]; a[root] = a[child]; a[child] = t;
        root = child;
      }
    }
    for (int end = n - 1; end > 0; end--) {
      int32_t t = a[0]; a[0] = a[end]; a[end] = t;
      int root = 0;
      for (;;) {
        int child = 2 * root + 1;
        if (child >= end) break;
        if (child + 1 < end && a[child] < a[child + 1]) child++;
        if (a[root] >= a[child]) break;
        int32_t t2 = a[root]; a[root] = a[child]; a[child] = t2;
        root = child;
      }
    }
  }
  }
}

__global__ void k_gather(RbxPoints P, RbxCells C) {
  RBX_SKIP_IF_VALID(C);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < P.n; k += gridDim.x * blockDim.x) {
    int g = C.gidx[k];
    C.sx[k] = P.x[g];
    C.sy[k] = P.y[g];
    C.sz[k] = P.z[g];
    C.sh[k] = P.h[g];
    C.sdem[k] = P.dem_id[g];
  }
}

// ---- parity mode: enumerate NNPS neighbours of dst points -----------------
__global__ void k_pairs(RbxPoints D, RbxCells C, double rs, int32_t *counts,
                        const int64_t *offsets, int32_t *idx) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= D.n) return;
  const RbxGridInfo gi = *C.info;
  int g = D.index ? D.index[k] : k;
  double px = D.x[g], py = D.y[g], pz = D.z[g], ph = D.h[g];
  double rs2 = rs * rs;
  double hi2 = rbx_h2(rs2, ph);
  double reach = gi.cell * (1.0 + 1e-9);
  int cx0 = rbx_cell_coord(px - reach, gi.x0, gi.inv_cell, gi.nx);
  int cx1 = rbx_cell_coord(px + reach, gi.x0, gi.inv_cell, gi.nx);
  int cy0 = rbx_cell_coord(py - reach, gi.y0, gi.inv_cell, gi.ny);
  int cy1 = rbx_cell_coord(py + reach, gi.y0, gi.inv_cell, gi.ny);
  int cz0 = rbx_cell_coord(pz - reach, gi.z0, gi.inv_cell, gi.nz);
  int cz1 = rbx_cell_coord(pz + reach, gi.z0, gi.inv_cell, gi.nz);
  int cnt = 0;
  int64_t base = idx ? offsets[k] : 0;
  for (int cz = cz0; cz <= cz1; cz++)
    for (int cy = cy0; cy <= cy1; cy++) {
      int row = (cz * gi.ny + cy) * gi.nx;
      int s = C.cell_start[row + cx0], e = C.cell_start[row + cx1 + 1];
      for (int q = s; q < e; q++) {
        double r2 = rbx_r2(px - C.sx[q], py - C.sy[q], pz - C.sz[q]);
        double hj2 = rbx_h2(rs2, C.sh[q]);
        if (r2 < hi2 || r2 < hj2) {
          if (idx) idx[base + cnt] = C.gidx[q];
          cnt++;
        }
      }
    }
  if (!idx) counts[k] = cnt;
}

}  // namespace

extern "C" size_t rbx_cells_workspace_bytes(int32_t cap_cells, int32_t cap_points) {
  size_t ntiles = ((size_t)cap_cells + 1 + kScanTile - 1) / kScanTile;
  size_t b = 256;                                   // BoundsWS, padded
  b += ((size_t)cap_cells + 1) * sizeof(int32_t);   // counts
  b += (ntiles + 1) * sizeof(int32_t);              // tile sums
  (void)cap_points;
  return (b + 255) & ~(size_t)255;
}

extern "C" int rbx_cells_build(const RbxPoints *pts, const RbxCells *cells, double min_cell,
                               uint32_t *status, void *workspace, size_t workspace_bytes,
                               void *stream_) {
  if (!pts || !cells || !workspace || !(min_cell > 0.)) return RBX_ERR_INVALID;
  if (pts->n > cells->cap_points || cells->cap_cells < 1) return RBX_ERR_INVALID;
  if (workspace_bytes < rbx_cells_workspace_bytes(cells->cap_cells, cells->cap_points))
    return RBX_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream_;
  char *w = (char *)workspace;
  BoundsWS *bws = (BoundsWS *)w;
  int32_t *counts = (int32_t *)(w + 256);
  int nscan = cells->cap_cells + 1;
  int ntiles = (nscan + kScanTile - 1) / kScanTile;
  int32_t *tile_sum = counts + nscan;

  cudaMemsetAsync(bws->mn, 0xFF, sizeof(bws->mn), st);
  cudaMemsetAsync(bws->mx, 0x00, sizeof(bws->mx) + 2 * sizeof(unsigned int), st);

  const int T = 256;
  int nb = rbx_blocks(pts->n, T);
  const int sms = rbx_sm_count();
  int bb = nb < sms * 8 ? nb : sms * 8;
  nb = nb < sms * 16 ? nb : sms * 16;     // grid-stride loops
  k_bounds<<<bb, T, 0, st>>>(*pts, *cells, min_cell, bws, status, counts, nscan);
  k_count<<<nb, T, 0, st>>>(*pts, *cells, counts);
  k_scan_tiles<<<ntiles, kScanThreads, 0, st>>>(*cells, counts, cells->cell_start, tile_sum, nscan);
  k_scan_sums<<<1, 1024, 0, st>>>(*cells, tile_sum, ntiles);
  k_scan_add<<<ntiles, kScanThreads, 0, st>>>(*cells, cells->cell_start, tile_sum, nscan);
  k_scatter<<<nb, T, 0, st>>>(*pts, *cells);
  {
    int ns = rbx_blocks(cells->cap_cells, T);
    k_sort_cells<<<ns < sms * 16 ? ns : sms * 16, T, 0, st>>>(*cells);
  }
  k_gather<<<nb, T, 0, st>>>(*pts, *cells);
  RBX_CHECK_LAUNCH();
  return RBX_OK;
}

extern "C" int rbx_pairs_dump(const RbxPoints *dst, const RbxCells *cells, double radius_scale,
                              int32_t *counts, const int64_t *offsets, int32_t *idx,
                              void *stream_) {
  if (!dst || !cells) return RBX_ERR_INVALID;
  if (!idx && !counts) return RBX_ERR_INVALID;
  if (idx && !offsets) return RBX_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream_;
  k_pairs<<<rbx_blocks(dst->n, 128), 128, 0, st>>>(*dst, *cells, radius_scale, counts, offsets, idx);
  RBX_CHECK_LAUNCH();
  return RBX_OK;
}
