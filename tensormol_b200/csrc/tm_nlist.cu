// K1: sorted cell-list build + neighbour enumeration with warp-ballot compaction (sm_100a).
//
// Replaces MolEmb.Make_NListNaive (C_API/MolEmb.cpp:1180-1247: x-sorted sweep on the host)
// and the Python pair loops of NeighborList.buildPairs (Neighbors.py:75-115).  The accept test
// is the reference's, bit for bit:  sqrt(dx*dx+dy*dy+dz*dz) + 1e-13 < rc  in float64 with one
// rounding per operation (no FMA contraction), MolEmb.cpp:1213-1218.
//
// Pipeline (all sizes bounded on the host, all geometry decided on the device):
//   bbox -> grid params -> per-slot cell id + rank (atomic) -> exclusive scan of cell counts
//   -> scatter -> per-cell sort by slot (determinism) + gather SAtom copy
//   -> centre rows sorted by (element, cell order) -> neighbour count -> scan -> fill.
#include "tm_internal.h"
#include <cstdio>
#include <algorithm>

#define FULL 0xffffffffu

// ---------------------------------------------------------------- ordered double <-> u64
__device__ __forceinline__ unsigned long long enc_d(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_d(unsigned long long b) {
  b = (b & 0x8000000000000000ull) ? (b & 0x7fffffffffffffffull) : ~b;
  return __longlong_as_double((long long)b);
}

// ---------------------------------------------------------------- tessellation (Periodic.py:131-168)
// slot = b*nreal + a ; block 0 = real atoms, then images for i,j,k in [-ntess..ntess]^3 skipping (0,0,0).
// coords_ + i*L0 + j*L1 + k*L2 is evaluated left to right with separate roundings like numpy does.
// ilo..ihi: image indices along the first lattice vector that can reach a slab rank's window (all of them otherwise);
// the other blocks are never binned (image_block_skipped below) and are not written.
// The lattice travels as a kernel argument (no pageable copy: the call sequence stays CUDA-graph capturable); the kernel
// also publishes 1/natom (L.v[9]) for the charge kernels.
__global__ void k_tessellate(const double* __restrict__ xyz, const int32_t* __restrict__ Z, int64_t nreal,
                             const __grid_constant__ LatArgs L, int ntess, int ilo, int ihi, double* __restrict__ pos, int32_t* __restrict__ Zo,
                             double* __restrict__ inv_n) {
  TM_PDL_PROLOGUE;
  const double* lat = L.v;
  if (blockIdx.x == 0 && threadIdx.x == 0) inv_n[0] = L.v[9];
  int side = 2 * ntess + 1;
  int64_t nimg = (int64_t)side * side * side;
  int64_t total = nimg * nreal;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = t / nreal, a = t - b * nreal;
    if (b == 0 && (0 < ilo || 0 > ihi)) continue;
    double x, y, z;
    if (b == 0) { x = xyz[3 * a]; y = xyz[3 * a + 1]; z = xyz[3 * a + 2]; }
    if (b > 0) {
      // block index b>0 enumerates (i,j,k) in loop order with the centre cell skipped
      int64_t centre = ((int64_t)ntess * side + ntess) * side + ntess;
      int64_t lin = (b - 1 < centre) ? (b - 1) : b;
      int k = (int)(lin % side) - ntess;
      int j = (int)((lin / side) % side) - ntess;
      int i = (int)(lin / ((int64_t)side * side)) - ntess;
      if (i < ilo || i > ihi) continue;
      x = xyz[3 * a]; y = xyz[3 * a + 1]; z = xyz[3 * a + 2];
      double di = (double)i, dj = (double)j, dk = (double)k;
      x = __dadd_rn(__dadd_rn(__dadd_rn(x, __dmul_rn(di, lat[0])), __dmul_rn(dj, lat[3])), __dmul_rn(dk, lat[6]));
      y = __dadd_rn(__dadd_rn(__dadd_rn(y, __dmul_rn(di, lat[1])), __dmul_rn(dj, lat[4])), __dmul_rn(dk, lat[7]));
      z = __dadd_rn(__dadd_rn(__dadd_rn(z, __dmul_rn(di, lat[2])), __dmul_rn(dj, lat[5])), __dmul_rn(dk, lat[8]));
    }
    pos[3 * t] = x;
    pos[3 * t + 1] = y;
    pos[3 * t + 2] = z;
    Zo[t] = Z[a];
  }
}

int tm_launch_tessellate(tm_ctx* c, const double* xyz_real, const int32_t* Z_real, int64_t nreal, const LatArgs& lat, int ntess, int ilo, int ihi) {
  int side = 2 * ntess + 1;
  int64_t total = (int64_t)side * side * side * nreal;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  TM_LAUNCH(k_tessellate, blocks, 256, 0, c->stream, xyz_real, Z_real, nreal, lat, ntess, ilo, ihi, (double*)c->b_pos.p, (int32_t*)c->b_Z.p, (double*)c->b_natom.p);
  c->launches++;
  TM_CUDA(cudaGetLastError());
  return TM_OK;
}

// ---------------------------------------------------------------- bbox
// Slab runs bin only the slots inside the owned slab plus the interaction halo (fractional coordinate window along the
// first lattice vector); everything else never enters the cell list of this rank.
struct Window { int on; double gx, gy, gz, lo, hi; int64_t nreal; int ntess, ilo, ihi; };
// image block of slot t along the first lattice vector outside [ilo, ihi]: the slot was not even tessellated
__device__ __forceinline__ bool image_block_skipped(int64_t t, const Window& w) {
  if (!w.on) return false;
  int64_t b = t / w.nreal;
  int i = 0;
  if (b > 0) {
    int side = 2 * w.ntess + 1;
    int64_t centre = ((int64_t)w.ntess * side + w.ntess) * side + w.ntess;
    int64_t lin = (b - 1 < centre) ? (b - 1) : b;
    i = (int)(lin / ((int64_t)side * side)) - w.ntess;
  }
  return i < w.ilo || i > w.ihi;
}
__device__ __forceinline__ bool slot_in_window(const double* __restrict__ pos, int64_t t, const Window& w) {
  if (!w.on) return true;
  if (image_block_skipped(t, w)) return false;
  double f = pos[3 * t] * w.gx + pos[3 * t + 1] * w.gy + pos[3 * t + 2] * w.gz;
  return f >= w.lo && f <= w.hi;
}
__global__ void k_bbox_init(unsigned long long* bb) {
  TM_PDL_PROLOGUE;
  if (threadIdx.x < 3) bb[threadIdx.x] = 0xffffffffffffffffull;   // mins
  else if (threadIdx.x < 6) bb[threadIdx.x] = 0ull;               // maxs
}

__global__ void k_bbox(const double* __restrict__ pos, const int32_t* __restrict__ Z, int64_t n, Window win, unsigned long long* bb) {
  TM_PDL_PROLOGUE;
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    if (image_block_skipped(t, win) || Z[t] <= 0 || !slot_in_window(pos, t, win)) continue;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      double v = pos[3 * t + d];
      mn[d] = fmin(mn[d], v);
      mx[d] = fmax(mx[d], v);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; d++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fmin(mn[d], __shfl_xor_sync(FULL, mn[d], o));
      mx[d] = fmax(mx[d], __shfl_xor_sync(FULL, mx[d], o));
    }
  }
  // block-level combine in shared memory, then 6 atomics per block (per-warp atomics on 6 addresses serialised: 42 us)
  __shared__ double smn[8][3], smx[8][3];
  int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; d++) { smn[w][d] = mn[d]; smx[w][d] = mx[d]; }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    int d = threadIdx.x;
    double a = smn[0][d], b = smx[0][d];
    for (int k = 1; k < (int)(blockDim.x >> 5); k++) { a = fmin(a, smn[k][d]); b = fmax(b, smx[k][d]); }
    if (a < 1e299) atomicMin(&bb[d], enc_d(a));
    if (b > -1e299) atomicMax(&bb[3 + d], enc_d(b));
  }
}

// one thread: grid geometry from the bbox; grows the cell edge until nmol*gx*gy*gz fits the cap
__global__ void k_grid_params(const unsigned long long* bb, double rc, int64_t nmol, int64_t ncells_cap, GridParams* g) {
  TM_PDL_PROLOGUE;
  double mn[3], mx[3];
  for (int d = 0; d < 3; d++) {
    mn[d] = dec_d(bb[d]);
    mx[d] = dec_d(bb[3 + d]);
    if (!(mx[d] >= mn[d])) { mn[d] = 0.0; mx[d] = 0.0; }   // empty input
  }
  double cell = rc * (1.0 + 1e-6);
  int gx, gy, gz;
  for (int it = 0; it < 200; it++) {
    gx = (int)floor((mx[0] - mn[0]) / cell) + 1;
    gy = (int)floor((mx[1] - mn[1]) / cell) + 1;
    gz = (int)floor((mx[2] - mn[2]) / cell) + 1;
    double tot = (double)gx * gy * gz * (double)nmol;
    if (tot <= (double)ncells_cap) break;
    cell *= 1.2599210498948732;
  }
  g->ox = mn[0]; g->oy = mn[1]; g->oz = mn[2];
  g->cell = cell;
  g->inv_cell = 1.0 / cell;
  g->gx = gx; g->gy = gy; g->gz = gz;
  g->zcell = cell; g->inv_zcell = 1.0 / cell; g->zdiv = 1;
  g->ncell_mol = gx * gy * gz;
  g->ncells = (int)(nmol * (int64_t)(gx * gy * gz));
}


// per slot: cell id and arrival rank inside the cell
// check_inside: the grid was laid out by the host from the lattice; an atom outside it means the caller did not wrap
// the coordinates into the cell (flag 8).  It is still binned (clamped), so nothing reads out of bounds.  In that
// case the grid arrives as the kernel argument `hg` and this kernel publishes it at *gp for the kernels that follow.
__global__ void k_cell_count(const double* __restrict__ pos, const int32_t* __restrict__ Z, int64_t n, int64_t maxnatom,
                             GridParams* gp, const __grid_constant__ GridParams hg, Window win, int check_inside, int32_t* __restrict__ cellid,
                             int32_t* __restrict__ rank, int32_t* __restrict__ count, int32_t* __restrict__ flags) {
  TM_PDL_PROLOGUE;
  GridParams g = check_inside ? hg : *gp;
  if (check_inside && blockIdx.x == 0 && threadIdx.x == 0) *gp = hg;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    if (image_block_skipped(t, win)) continue;   // k_scatter applies the same test: cellid / rank of these slots are never read
    if (Z[t] <= 0 || !slot_in_window(pos, t, win)) { cellid[t] = -1; rank[t] = -1; continue; }   // rank[] is reused as sidx_of_slot (-1 = absent)
    if (check_inside) {
      double fx = (pos[3 * t] - g.ox) * g.inv_cell, fy = (pos[3 * t + 1] - g.oy) * g.inv_cell, fz = (pos[3 * t + 2] - g.oz) * g.inv_zcell;
      if (fx < 0.0 || fy < 0.0 || fz < 0.0 || fx > (double)g.gx || fy > (double)g.gy || fz > (double)g.gz) atomicOr(flags, 8);
    }
    int cx = cell_coord(pos[3 * t], g.ox, g.inv_cell, g.gx);
    int cy = cell_coord(pos[3 * t + 1], g.oy, g.inv_cell, g.gy);
    int cz = cell_coord(pos[3 * t + 2], g.oz, g.inv_zcell, g.gz);
    int m = (int)(t / maxnatom);
    int cid = m * g.ncell_mol + (cx * g.gy + cy) * g.gz + cz;   // z fastest: a +-1 z-run is contiguous
    cellid[t] = cid;
    rank[t] = atomicAdd(&count[cid], 1);
  }
}

// ---------------------------------------------------------------- exclusive scan (3 kernels, int32)
#define SCAN_TILE 2048   // elements per block (256 threads x 8)
__global__ void k_scan_local(const int32_t* __restrict__ in, int32_t* __restrict__ out, int32_t* __restrict__ blksum, int64_t n) {
  TM_PDL_PROLOGUE;
  __shared__ int32_t wsum[8];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * 8;
  int32_t v[8], s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t t = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  int32_t woff = 0;
  for (int i = 0; i < w; i++) woff += wsum[i];
  int32_t run = woff + inc - s;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
  if (threadIdx.x == 255) blksum[blockIdx.x] = woff + inc;
}
// single block: exclusive scan of block sums (nblk <= 1<<20), also writes the grand total at [nblk]
__global__ void k_scan_blocks(int32_t* blksum, int nblk) {
  TM_PDL_PROLOGUE;
  __shared__ int32_t wsum[32];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < nblk; base += 1024) {
    int i = base + threadIdx.x;
    int32_t v = (i < nblk) ? blksum[i] : 0;
    int32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(FULL, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
      int32_t x = wsum[lane], xi = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int32_t t = __shfl_up_sync(FULL, xi, o);
        if (lane >= o) xi += t;
      }
      wsum[lane] = xi - x;   // exclusive warp offsets
    }
    __syncthreads();
    int32_t excl = carry + wsum[w] + inc - v;
    if (i < nblk) blksum[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) blksum[nblk] = carry;
}
__global__ void k_scan_add(int32_t* __restrict__ out, const int32_t* __restrict__ blksum, int64_t n, int nblk) {
  TM_PDL_PROLOGUE;
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * 8;
  int32_t off = blksum[blockIdx.x];
#pragma unroll
  for (int i = 0; i < 8; i++)
    if (base + i < n) out[base + i] += off;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = blksum[nblk];   // total at out[n]
}

// k_scan_blocks + k_scan_add in one launch for moderate block counts: every block sums the (unscanned) totals of the
// blocks before it by itself
__global__ void k_scan_add_self(int32_t* __restrict__ out, const int32_t* __restrict__ blksum, int64_t n, int nblk) {
  TM_PDL_PROLOGUE;
  __shared__ int32_t wsum[8];
  __shared__ int32_t off_s;
  int32_t part = 0;
  const int upto = (blockIdx.x == 0) ? nblk : (int)blockIdx.x;   // block 0 also needs the grand total
  for (int i = threadIdx.x; i < upto; i += 256) part += blksum[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t t = 0;
    for (int i = 0; i < 8; i++) t += wsum[i];
    if (blockIdx.x == 0) { out[n] = t; t = 0; }   // total at out[n]
    off_s = t;
  }
  __syncthreads();
  int32_t off = off_s;
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * 8;
#pragma unroll
  for (int i = 0; i < 8; i++)
    if (base + i < n) out[base + i] += off;
}

// exclusive scan of in[0..n) into out[0..n], out[n] = total.  tmp needs (nblk+1) ints.
static int scan_exclusive(tm_ctx* c, const int32_t* in, int32_t* out, int64_t n, int32_t* tmp) {
  int nblk = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
  if (nblk < 1) nblk = 1;
  TM_LAUNCH(k_scan_local, nblk, 256, 0, c->stream, in, out, tmp, n);
  if (nblk <= 2048) {
    TM_LAUNCH(k_scan_add_self, nblk, 256, 0, c->stream, out, tmp, n, nblk);
    c->launches += 2;
    TM_CUDA(cudaGetLastError());
    return TM_OK;
  }
  TM_LAUNCH(k_scan_blocks, 1, 1024, 0, c->stream, tmp, nblk);
  TM_LAUNCH(k_scan_add, nblk, 256, 0, c->stream, out, tmp, n, nblk);
  c->launches += 3;
  TM_CUDA(cudaGetLastError());
  return TM_OK;
}

__global__ void k_scatter(const int32_t* __restrict__ cellid, const int32_t* __restrict__ rank, const int32_t* __restrict__ cstart,
                          int64_t n, Window win, int32_t* __restrict__ sorted) {
  TM_PDL_PROLOGUE;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    if (image_block_skipped(t, win)) continue;
    int cid = cellid[t];
    if (cid >= 0) sorted[cstart[cid] + rank[t]] = (int32_t)t;
  }
}

// One warp per cell: order the cell's slots ascending (removes the atomic-arrival nondeterminism) by rank counting
// across lanes, then write the 32-byte SAtom records in sorted order (lane = record: coalesced).  Cells holding more
// than 32 atoms fall back to a serial insertion sort by lane 0.
__global__ void k_cell_sort_gather(const GridParams* __restrict__ gp, const int32_t* __restrict__ cstart, int32_t* __restrict__ sorted,
                                   const double* __restrict__ pos, const int32_t* __restrict__ Z, const __grid_constant__ DevParams P,
                                   SAtom* __restrict__ sat, int32_t* __restrict__ sidx_of_slot) {
  TM_PDL_PROLOGUE;
  int ncells = gp->ncells;
  int lane = threadIdx.x & 31;
  int warps = (gridDim.x * blockDim.x) >> 5;
  for (int cid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; cid < ncells; cid += warps) {
    int b = cstart[cid], e = cstart[cid + 1], n = e - b;
    if (n <= 0) continue;
    if (n <= 32) {
      int32_t mine = (lane < n) ? sorted[b + lane] : 0x7fffffff;
      int rank = 0;
      for (int k = 0; k < n; k++) {
        int32_t other = __shfl_sync(FULL, mine, k);
        rank += (other < mine) ? 1 : 0;
      }
      __syncwarp();
      if (lane < n) sorted[b + rank] = mine;
      __syncwarp();
    } else if (lane == 0) {
      for (int i = b + 1; i < e; i++) {
        int32_t v = sorted[i];
        int j = i - 1;
        while (j >= b && sorted[j] > v) { sorted[j + 1] = sorted[j]; j--; }
        sorted[j + 1] = v;
      }
    }
    __syncwarp();
    for (int i = b + lane; i < e; i += 32) {
      int32_t s = sorted[i];
      SAtom a;
      a.x = pos[3 * (int64_t)s]; a.y = pos[3 * (int64_t)s + 1]; a.z = pos[3 * (int64_t)s + 2];
      a.slot = s;
      int z = Z[s], ei = -1;
      for (int k = 0; k < P.n_ele; k++) if (P.eles[k] == z) ei = k;
      a.e = ei;
      sat[i] = a;
      sidx_of_slot[s] = i;
    }
  }
}

// ---------------------------------------------------------------- build launcher
int tm_launch_nlist_build(tm_ctx* c, const SysView& s, double rc_grid) {
  int64_t n = s.nslots;
  int rc;
  if ((rc = tm_buf(c, c->b_cellid, n * 4))) return rc;
  if ((rc = tm_buf(c, c->b_rank, n * 4))) return rc;
  if ((rc = tm_buf(c, c->b_count, (s.ncells_cap + 8) * 4))) return rc;
  if ((rc = tm_buf(c, c->b_cstart, (s.ncells_cap + 8) * 4))) return rc;
  if ((rc = tm_buf(c, c->b_sorted, n * 4))) return rc;
  if ((rc = tm_buf(c, c->b_satom, n * sizeof(SAtom)))) return rc;
  if ((rc = tm_buf(c, c->b_rowofslot, n * 4))) return rc;   // reused as sidx_of_slot first (see rows)
  if ((rc = tm_buf(c, c->b_scan_tmp, ((s.ncells_cap + n) / SCAN_TILE + 16) * 4 * 2))) return rc;
  if ((rc = tm_buf(c, c->b_bbox, 64))) return rc;
  if ((rc = tm_buf(c, c->b_grid, sizeof(GridParams)))) return rc;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  unsigned long long* bb = (unsigned long long*)c->b_bbox.p;
  GridParams* gp = (GridParams*)c->b_grid.p;
  Window win{s.window_on, s.slab_g[0], s.slab_g[1], s.slab_g[2], s.win_lo, s.win_hi, s.nreal > 0 ? s.nreal : 1, s.win_ntess, s.win_ilo, s.win_ihi};
  int64_t ncs = s.ncells_cap;   // cells the count / scan passes have to cover
  if (s.grid_host) {
    ncs = s.hgrid.ncells;     // k_cell_count publishes the grid
  } else {
    TM_LAUNCH(k_bbox_init, 1, 32, 0, c->stream, bb);
    TM_LAUNCH(k_bbox, blocks, 256, 0, c->stream, (const double*)c->b_pos.p, (const int32_t*)c->b_Z.p, n, win, bb);
    TM_LAUNCH(k_grid_params, 1, 1, 0, c->stream, bb, rc_grid, s.nmol, s.ncells_cap, gp);
    c->launches += 3;
  }
  TM_CUDA(cudaMemsetAsync(c->b_count.p, 0, (ncs + 8) * 4, c->stream));
  TM_LAUNCH(k_cell_count, blocks, 256, 0, c->stream, (const double*)c->b_pos.p, (const int32_t*)c->b_Z.p, n, s.maxnatom, gp, s.hgrid, win, s.grid_host,
                                              (int32_t*)c->b_cellid.p, (int32_t*)c->b_rank.p, (int32_t*)c->b_count.p, (int32_t*)c->b_flags.p);
  c->launches += 1;
  if ((rc = scan_exclusive(c, (const int32_t*)c->b_count.p, (int32_t*)c->b_cstart.p, ncs, (int32_t*)c->b_scan_tmp.p))) return rc;
  TM_LAUNCH(k_scatter, blocks, 256, 0, c->stream, (const int32_t*)c->b_cellid.p, (const int32_t*)c->b_rank.p, (const int32_t*)c->b_cstart.p, n, win,
                                           (int32_t*)c->b_sorted.p);
  int cblocks = (int)((ncs * 32 + 255) / 256);
  if (cblocks > 148 * 32) cblocks = 148 * 32;
  TM_LAUNCH(k_cell_sort_gather, cblocks, 256, 0, c->stream, gp, (const int32_t*)c->b_cstart.p, (int32_t*)c->b_sorted.p, (const double*)c->b_pos.p,
                                                     (const int32_t*)c->b_Z.p, c->hp, (SAtom*)c->b_satom.p, (int32_t*)c->b_rank.p);
  c->launches += 2;
  TM_CUDA(cudaGetLastError());
  return TM_OK;
}

// ---------------------------------------------------------------- centre rows
// Rows are the centres ordered by (element, cell-sorted position); each element's row range starts
// at a multiple of TM_ROW_TILE so a GEMM row tile never straddles two elements.
//   rowmeta[2e] = first row of element e, rowmeta[2e+1] = number of rows of element e,
//   rowmeta[2*TM_MAX_ELE] = total centres.
#define ROWS_BLOCK 1024
struct SlabFilter { int rank, world; double gx, gy, gz; };
// Owned-atom slabs: a centre belongs to rank floor(frac*world) where frac = pos . g is its fractional
// coordinate along the first lattice vector (g = first row of the inverse lattice, supplied by the host).
__device__ __forceinline__ bool is_centre(const SAtom& a, int64_t nreal, int periodic, const SlabFilter& sf) {
  if (a.e < 0) return false;
  if (periodic && a.slot >= nreal) return false;
  if (sf.world > 1) {
    double frac = a.x * sf.gx + a.y * sf.gy + a.z * sf.gz;
    int owner = (int)floor(frac * (double)sf.world);
    owner = owner < 0 ? 0 : (owner >= sf.world ? sf.world - 1 : owner);
    if (owner != sf.rank) return false;
  }
  return true;
}

// Also resets the row tables that k_rows_fill (the launch after next) fills in: rowslot[0..nrows) = rowofslot[0..nq) = -1.
__global__ void k_rows_count(const SAtom* __restrict__ sat, const int32_t* __restrict__ cstart, const GridParams* __restrict__ gp,
                             int64_t nreal, int periodic, SlabFilter sf, int32_t* __restrict__ blkcnt, int32_t* __restrict__ rowslot,
                             int64_t nrows, int32_t* __restrict__ rowofslot, int64_t nq) {
  TM_PDL_PROLOGUE;
  __shared__ int32_t cnt[TM_MAX_ELE];
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nrows + nq; t += (int64_t)gridDim.x * blockDim.x) {
    if (t < nrows) rowslot[t] = -1;
    else rowofslot[t - nrows] = -1;
  }
  if (threadIdx.x < TM_MAX_ELE) cnt[threadIdx.x] = 0;
  __syncthreads();
  GridParams g = *gp;
  int ntot = cstart[g.ncells];
  if (blockIdx.x * ROWS_BLOCK >= ntot) return;   // k_rows_scan only visits the blocks below ntot
  int i = blockIdx.x * ROWS_BLOCK + threadIdx.x;
  if (i < ntot) {
    SAtom a = sat[i];
    if (is_centre(a, nreal, periodic, sf)) atomicAdd(&cnt[a.e], 1);
  }
  __syncthreads();
  if (threadIdx.x < TM_MAX_ELE) blkcnt[blockIdx.x * TM_MAX_ELE + threadIdx.x] = cnt[threadIdx.x];
}

// single block: per element exclusive scan over blocks; element bases padded to TM_ROW_TILE
__global__ void k_rows_scan(int32_t* __restrict__ blkcnt, int nblk, int n_ele, const int32_t* __restrict__ cstart,
                            const GridParams* __restrict__ gp, int32_t* __restrict__ rowmeta) {
  TM_PDL_PROLOGUE;
  __shared__ int32_t tot[TM_MAX_ELE];
  nblk = min(nblk, (cstart[gp->ncells] + ROWS_BLOCK - 1) / ROWS_BLOCK);
  int e = threadIdx.x >> 5, lane = threadIdx.x & 31;   // one warp per element; a lane owns 32 consecutive blocks of a 1024-block chunk
  if (e < TM_MAX_ELE) {
    int32_t carry = 0;
    for (int base = 0; base < nblk; base += 1024) {
      int32_t v[32];
      int32_t sum = 0;
#pragma unroll
      for (int k = 0; k < 32; k++) {   // all 32 loads in flight (the serial version waited one L2 round trip per block)
        int i = base + lane * 32 + k;
        v[k] = (i < nblk) ? blkcnt[i * TM_MAX_ELE + e] : 0;
        sum += v[k];
      }
      int32_t inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int32_t t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
      }
      int32_t run = carry + inc - sum;
#pragma unroll
      for (int k = 0; k < 32; k++) {
        int i = base + lane * 32 + k;
        if (i < nblk) blkcnt[i * TM_MAX_ELE + e] = run;
        run += v[k];
      }
      carry += __shfl_sync(FULL, inc, 31);
    }
    if (lane == 0) tot[e] = carry;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t base = 0, total = 0;
    for (int k = 0; k < TM_MAX_ELE; k++) {
      int32_t cnt = (k < n_ele) ? tot[k] : 0;
      rowmeta[2 * k] = base;
      rowmeta[2 * k + 1] = cnt;
      base += ((cnt + TM_ROW_TILE - 1) / TM_ROW_TILE) * TM_ROW_TILE;
      total += cnt;
    }
    rowmeta[2 * TM_MAX_ELE] = total;
    rowmeta[2 * TM_MAX_ELE + 1] = base;   // rows in use incl. padding
  }
}

// self_scan != 0 (moderate block counts): blkcnt holds the raw per-block counts of k_rows_count and every block forms
// its own prefix and the element totals from them (k_rows_scan is not launched; block 0 writes rowmeta).
__global__ void k_rows_fill(const SAtom* __restrict__ sat, const int32_t* __restrict__ cstart, const GridParams* __restrict__ gp,
                            int64_t nreal, int periodic, SlabFilter sf, const int32_t* __restrict__ blkcnt,
                            int32_t* __restrict__ rowmeta, int32_t* __restrict__ rowslot, int32_t* __restrict__ rowsidx,
                            int32_t* __restrict__ rowofslot, int self_scan, int nblk, int n_ele) {
  TM_PDL_PROLOGUE;
  __shared__ int32_t wcnt[32][TM_MAX_ELE];
  __shared__ int32_t s_pre[32][TM_MAX_ELE], s_tot[32][TM_MAX_ELE];
  __shared__ int32_t s_base[TM_MAX_ELE], s_mine[TM_MAX_ELE];
  GridParams g = *gp;
  int ntot = cstart[g.ncells];
  if (blockIdx.x * ROWS_BLOCK >= ntot && !(self_scan && blockIdx.x == 0)) return;
  if (self_scan) {
    static_assert(TM_MAX_ELE == 8, "lane & 7 selects the element");
    const int live = min(nblk, (ntot + ROWS_BLOCK - 1) / ROWS_BLOCK);   // blocks that counted anything
    const int el = threadIdx.x & 7;
    int32_t pre = 0, tot = 0;
    for (int i = threadIdx.x >> 3; i < live; i += ROWS_BLOCK / 8) {
      int32_t v = blkcnt[i * TM_MAX_ELE + el];
      tot += v;
      if (i < (int)blockIdx.x) pre += v;
    }
    pre += __shfl_xor_sync(FULL, pre, 8); tot += __shfl_xor_sync(FULL, tot, 8);
    pre += __shfl_xor_sync(FULL, pre, 16); tot += __shfl_xor_sync(FULL, tot, 16);
    if ((threadIdx.x & 31) < 8) { s_pre[threadIdx.x >> 5][el] = pre; s_tot[threadIdx.x >> 5][el] = tot; }
    __syncthreads();
    if (threadIdx.x < TM_MAX_ELE) {
      int32_t p2 = 0, t2 = 0;
      for (int w = 0; w < 32; w++) { p2 += s_pre[w][threadIdx.x]; t2 += s_tot[w][threadIdx.x]; }
      s_mine[threadIdx.x] = p2;
      s_tot[0][threadIdx.x] = (threadIdx.x < n_ele) ? t2 : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int32_t base = 0, total = 0;
      for (int k = 0; k < TM_MAX_ELE; k++) {
        int32_t cnt = s_tot[0][k];
        s_base[k] = base;
        if (blockIdx.x == 0) { rowmeta[2 * k] = base; rowmeta[2 * k + 1] = cnt; }
        base += ((cnt + TM_ROW_TILE - 1) / TM_ROW_TILE) * TM_ROW_TILE;
        total += cnt;
      }
      if (blockIdx.x == 0) { rowmeta[2 * TM_MAX_ELE] = total; rowmeta[2 * TM_MAX_ELE + 1] = base; }
    }
    __syncthreads();
    if (blockIdx.x * ROWS_BLOCK >= ntot) return;
  }
  int i = blockIdx.x * ROWS_BLOCK + threadIdx.x;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int e = -1, slot = -1;
  if (i < ntot) {
    SAtom a = sat[i];
    slot = a.slot;
    if (is_centre(a, nreal, periodic, sf)) e = a.e;
  }
  // rank of this thread among same-element centres of the block, in index order
  int myrank = 0;
#pragma unroll
  for (int k = 0; k < TM_MAX_ELE; k++) {
    unsigned m = __ballot_sync(FULL, e == k);
    if (e == k) myrank = __popc(m & ((1u << lane) - 1));
    if (lane == 0) wcnt[w][k] = __popc(m);
  }
  __syncthreads();
  if (e >= 0) {
    int off = 0;
    for (int ww = 0; ww < w; ww++) off += wcnt[ww][e];
    int row = (self_scan ? s_base[e] + s_mine[e] : rowmeta[2 * e] + blkcnt[blockIdx.x * TM_MAX_ELE + e]) + off + myrank;
    rowslot[row] = slot;
    rowsidx[row] = i;
    rowofslot[slot] = row;
  }
}

int tm_launch_rows(tm_ctx* c, const SysView& s) {
  int rc;
  int nblk = (int)((s.nslots + ROWS_BLOCK - 1) / ROWS_BLOCK);
  if (nblk < 1) nblk = 1;
  if ((rc = tm_buf(c, c->b_blkcnt, (size_t)nblk * TM_MAX_ELE * 4))) return rc;
  if ((rc = tm_buf(c, c->b_rowmeta, (2 * TM_MAX_ELE + 2) * 4))) return rc;
  if ((rc = tm_buf(c, c->b_rowslot, s.nrows * 4))) return rc;
  if ((rc = tm_buf(c, c->b_rowsidx, s.nrows * 4))) return rc;
  // rowofslot is only read for slots that can be centres (the descriptor output), i.e. the real block in images mode
  int64_t nq = s.periodic ? s.nreal : s.nslots;
  const SAtom* sat = (const SAtom*)c->b_satom.p;
  const GridParams* gp = (const GridParams*)c->b_grid.p;
  SlabFilter sf{s.slab_rank, s.slab_world, s.slab_g[0], s.slab_g[1], s.slab_g[2]};
  TM_LAUNCH(k_rows_count, nblk, ROWS_BLOCK, 0, c->stream, sat, (const int32_t*)c->b_cstart.p, gp, s.nreal, s.periodic, sf,
                                                   (int32_t*)c->b_blkcnt.p, (int32_t*)c->b_rowslot.p, s.nrows, (int32_t*)c->b_rowofslot.p, nq);
  const int self_scan = nblk <= 2048 ? 1 : 0;   // every fill block re-reads the nblk counts: only while that is cheap
  if (!self_scan) {
    TM_LAUNCH(k_rows_scan, 1, 32 * TM_MAX_ELE, 0, c->stream, (int32_t*)c->b_blkcnt.p, nblk, c->hp.n_ele, (const int32_t*)c->b_cstart.p, gp,
                                                      (int32_t*)c->b_rowmeta.p);
    c->launches++;
  }
  TM_LAUNCH(k_rows_fill, nblk, ROWS_BLOCK, 0, c->stream, sat, (const int32_t*)c->b_cstart.p, gp, s.nreal, s.periodic, sf,
                                                  (const int32_t*)c->b_blkcnt.p, (int32_t*)c->b_rowmeta.p, (int32_t*)c->b_rowslot.p,
                                                  (int32_t*)c->b_rowsidx.p, (int32_t*)c->b_rowofslot.p, self_scan, nblk, c->hp.n_ele);
  c->launches += 2;
  TM_CUDA(cudaGetLastError());
  return TM_OK;
}

// ---------------------------------------------------------------- neighbour enumeration
// One warp per centre row.  The 27 neighbour cells are visited as 9 (x,y) columns whose +-1 z-run
// is contiguous in the cell-sorted array; lanes stride over the run (coalesced 32-byte records),
// the accept test runs in float64, survivors are compacted with __ballot_sync / __popc.
__device__ __forceinline__ double ref_dist(const SAtom& a, double xi, double yi, double zi) {
  double dx = __dsub_rn(xi, a.x), dy = __dsub_rn(yi, a.y), dz = __dsub_rn(zi, a.z);
  double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  return __dadd_rn(__dsqrt_rn(d2), 0.0000000000001);
}

// Entry = cell-sorted index of j, bit 31 set when j is also inside the angular cutoff.
// One pass: row r owns the TM_NB_STRIDE slots nbr[r*TM_NB_STRIDE ...]; nbcnt[r] = number written.  More than
// TM_NB_STRIDE radial neighbours of one centre raises flag 2 (TM_ECAP).
// The accept test `sqrt(d2) + 1e-13 < rc` (MolEmb.cpp:1213-1218, one rounding per operation) decided from d2 alone
// wherever that is safe: surely inside below (rc - 2e-13)^2 (1 - 1e-15), surely outside above rc^2 (1 + 1e-15); only the
// band in between (relative width ~1e-13) takes the exact square-root path.
struct AcceptBand { double lo, hi, rc; };
__device__ __forceinline__ AcceptBand accept_band(double rc) {
  AcceptBand a;
  double t = rc - 2.0e-13;
  a.lo = t > 0.0 ? t * t * (1.0 - 1.0e-15) : -1.0;
  a.hi = rc * rc * (1.0 + 1.0e-15);
  a.rc = rc;
  return a;
}
__device__ __forceinline__ bool accept(double d2, const AcceptBand& a) {
  if (d2 < a.lo) return true;
  if (d2 > a.hi) return false;
  return __dadd_rn(__dsqrt_rn(d2), 0.0000000000001) < a.rc;
}

// The nine (x, y) columns of a centre are resolved by nine lanes at once and their z-runs walked as ONE index space,
// three 32-candidate passes at a time with the loads of all three issued first (the column-by-column version was a
// chain of ~18 dependent L2 round trips per warp).  Order of a row = column order, then cell-sorted index: unchanged.
__global__ void k_neighbours(const SAtom* __restrict__ sat, const int32_t* __restrict__ cstart, const GridParams* __restrict__ gp,
                             const int32_t* __restrict__ rowsidx, const int32_t* __restrict__ rowslot, int64_t nrows, int64_t maxnatom,
                             double rr, double ra, int32_t* __restrict__ nbcnt, uint32_t* __restrict__ nbr, int32_t* __restrict__ flags) {
  TM_PDL_PROLOGUE;
  int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  int slot = rowslot[row];
  if (slot < 0) { if (lane == 0) nbcnt[row] = 0; return; }
  GridParams g = *gp;
  int si = rowsidx[row];
  SAtom ci = sat[si];
  int m = (int)(slot / maxnatom);
  int cx = cell_coord(ci.x, g.ox, g.inv_cell, g.gx);
  int cy = cell_coord(ci.y, g.oy, g.inv_cell, g.gy);
  // z bins that can hold a neighbour, from the centre's position (the bins may be finer than the cell edge)
  int z0 = cell_coord(ci.z - rr * (1.0 + 1e-9), g.oz, g.inv_zcell, g.gz), z1 = cell_coord(ci.z + rr * (1.0 + 1e-9), g.oz, g.inv_zcell, g.gz);
  // lane c < 9 owns column (dx, dy) = (c / 3 - 1, c % 3 - 1): run [cb, ce)
  int cb = 0, ce = 0;
  if (lane < 9) {
    int x = cx + lane / 3 - 1, y = cy + lane % 3 - 1;
    if (x >= 0 && x < g.gx && y >= 0 && y < g.gy) {
      int cbase = m * g.ncell_mol + (x * g.gy + y) * g.gz;
      cb = cstart[cbase + z0];
      ce = cstart[cbase + z1 + 1];
    }
  }
  int len = ce - cb, inc = len;
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) {
    int t = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += t;
  }
  const int T = __shfl_sync(FULL, inc, 8);
  // start[c] = first flat index of column c, off[c] = cb_c - start_c  (j = idx + off[c])
  int start[9], off[9];
#pragma unroll
  for (int c = 0; c < 9; c++) {
    start[c] = __shfl_sync(FULL, inc - len, c);
    off[c] = __shfl_sync(FULL, cb - (inc - len), c);
  }
  const AcceptBand br = accept_band(rr), ba = accept_band(ra);
  int total = 0;
  uint32_t* out = nbr + row * TM_NB_STRIDE;
  for (int i0 = 0; i0 < T; i0 += 96) {
    int jj[3];
    bool live[3];
    SAtom aj[3];
#pragma unroll
    for (int u = 0; u < 3; u++) {
      int idx = i0 + 32 * u + lane;
      int j = idx + off[0];
#pragma unroll
      for (int c = 1; c < 9; c++) j = (idx >= start[c]) ? idx + off[c] : j;
      jj[u] = j;
      live[u] = idx < T && j != si;
      if (live[u]) aj[u] = sat[j];
    }
#pragma unroll
    for (int u = 0; u < 3; u++) {
      if (i0 + 32 * u >= T) break;
      bool ok = false, ang = false;
      if (live[u]) {
        double dx = __dsub_rn(ci.x, aj[u].x), dy = __dsub_rn(ci.y, aj[u].y), dz = __dsub_rn(ci.z, aj[u].z);
        double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        ok = accept(d2, br);
        ang = ok && accept(d2, ba);
      }
      unsigned mk = __ballot_sync(FULL, ok);
      if (ok) {
        int w = total + __popc(mk & ((1u << lane) - 1));
        if (w < TM_NB_STRIDE) out[w] = (uint32_t)jj[u] | (ang ? 0x80000000u : 0u);
      }
      total += __popc(mk);
    }
  }
  if (lane == 0) {
    if (total > TM_NB_STRIDE) { atomicOr(flags, 2); total = TM_NB_STRIDE; }
    nbcnt[row] = total;
  }
}

int tm_launch_neighbours(tm_ctx* c, const SysView& s) {
  int rc;
  if ((rc = tm_buf(c, c->b_nbcnt, (s.nrows + 8) * 4))) return rc;
  if ((rc = tm_buf(c, c->b_nbr, (size_t)s.nrows * TM_NB_STRIDE * 4))) return rc;
  int blocks = (int)((s.nrows * 32 + 255) / 256);
  TM_LAUNCH(k_neighbours, blocks, 256, 0, c->stream, (const SAtom*)c->b_satom.p, (const int32_t*)c->b_cstart.p, (const GridParams*)c->b_grid.p,
                                              (const int32_t*)c->b_rowsidx.p, (const int32_t*)c->b_rowslot.p, s.nrows, s.maxnatom,
                                              c->hp.rr_exact + c->skin, c->hp.ra_exact + c->skin, (int32_t*)c->b_nbcnt.p, (uint32_t*)c->b_nbr.p,
                                              (int32_t*)c->b_flags.p);
  c->launches++;
  TM_CUDA(cudaGetLastError());
  return TM_OK;
}

// ---------------------------------------------------------------- MolEmb-compatible CSR list
// Rows are slots i < nreal (all slots when nreal == n).  do_perms == 0 keeps only j > i.
template <bool FILL>
__global__ void k_nlist_csr(const SAtom* __restrict__ sat, const int32_t* __restrict__ cstart, const GridParams* __restrict__ gp,
                            const int32_t* __restrict__ sidx_of_slot, int64_t ncentres, int64_t maxnatom, double rc, int do_perms,
                            int32_t* __restrict__ cnt, const int64_t* __restrict__ off, int32_t* __restrict__ out) {
  TM_PDL_PROLOGUE;
  int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (i >= ncentres) return;
  int si = sidx_of_slot[i];
  if (si < 0) { if (!FILL && lane == 0) cnt[i] = 0; return; }
  GridParams g = *gp;
  SAtom ci = sat[si];
  int m = (int)(i / maxnatom);
  int cx = cell_coord(ci.x, g.ox, g.inv_cell, g.gx);
  int cy = cell_coord(ci.y, g.oy, g.inv_cell, g.gy);
  int z0 = cell_coord(ci.z - rc * (1.0 + 1e-9), g.oz, g.inv_zcell, g.gz), z1 = cell_coord(ci.z + rc * (1.0 + 1e-9), g.oz, g.inv_zcell, g.gz);
  int total = 0;
  int64_t wbase = FILL ? off[i] : 0;
  for (int dx = -1; dx <= 1; dx++) {
    int x = cx + dx;
    if (x < 0 || x >= g.gx) continue;
    for (int dy = -1; dy <= 1; dy++) {
      int y = cy + dy;
      if (y < 0 || y >= g.gy) continue;
      int cbase = m * g.ncell_mol + (x * g.gy + y) * g.gz;
      int b = cstart[cbase + z0], e = cstart[cbase + z1 + 1];
      for (int j0 = b; j0 < e; j0 += 32) {
        int j = j0 + lane;
        bool ok = false;
        int js = -1;
        if (j < e && j != si) {
          SAtom aj = sat[j];
          js = aj.slot;
          ok = ref_dist(aj, ci.x, ci.y, ci.z) < rc;
          if (!do_perms && js < (int)i) ok = false;
        }
        unsigned mk = __ballot_sync(FULL, ok);
        if (FILL && ok) out[wbase + total + __popc(mk & ((1u << lane) - 1))] = js;
        total += __popc(mk);
      }
    }
  }
  if (!FILL && lane == 0) cnt[i] = total;
}

__global__ void k_mark_unsorted(int32_t* sidx_of_slot, int64_t n) {
  TM_PDL_PROLOGUE;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) sidx_of_slot[t] = -1;
}

// Builds the grid for `rc`, counts, and returns the CSR on the host in c->h_off / c->h_idx.
int tm_launch_nlist_csr(tm_ctx* c, const SysView& s, double rc, int do_perms, int64_t* total_out) {
  int r;
  // sidx_of_slot lives in b_rank after the build; mark invalid slots first
  if ((r = tm_launch_nlist_build(c, s, rc))) return r;
  int64_t ncent = s.periodic ? s.nreal : s.nslots;
  if ((r = tm_buf(c, c->b_nbcnt, (ncent + 8) * 4))) return r;
  const SAtom* sat = (const SAtom*)c->b_satom.p;
  const GridParams* gp = (const GridParams*)c->b_grid.p;
  int blocks = (int)((ncent * 32 + 255) / 256);
  if (blocks < 1) blocks = 1;
  TM_LAUNCH(k_nlist_csr<false>, blocks, 256, 0, c->stream, sat, (const int32_t*)c->b_cstart.p, gp, (const int32_t*)c->b_rank.p, ncent, s.maxnatom, rc,
                                                    do_perms, (int32_t*)c->b_nbcnt.p, nullptr, nullptr);
  c->launches++;
  TM_CUDA(cudaGetLastError());
  std::vector<int32_t> hc((size_t)ncent);
  TM_CUDA(cudaMemcpyAsync(hc.data(), c->b_nbcnt.p, (size_t)ncent * 4, cudaMemcpyDeviceToHost, c->stream));
  TM_CUDA(cudaStreamSynchronize(c->stream));
  c->h_off.assign((size_t)ncent + 1, 0);
  for (int64_t i = 0; i < ncent; i++) c->h_off[i + 1] = c->h_off[i] + hc[i];
  int64_t total = c->h_off[ncent];
  *total_out = total;
  if ((r = tm_buf(c, c->b_nboff, (ncent + 8) * 8))) return r;
  if ((r = tm_buf(c, c->b_nbr, (size_t)(total + 8) * 4))) return r;
  TM_CUDA(cudaMemcpyAsync(c->b_nboff.p, c->h_off.data(), (size_t)(ncent + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  TM_LAUNCH(k_nlist_csr<true>, blocks, 256, 0, c->stream, sat, (const int32_t*)c->b_cstart.p, gp, (const int32_t*)c->b_rank.p, ncent, s.maxnatom, rc,
                                                   do_perms, nullptr, (const int64_t*)c->b_nboff.p, (int32_t*)c->b_nbr.p);
  c->launches++;
  TM_CUDA(cudaGetLastError());
  std::vector<int32_t> hi((size_t)total);
  if (total) TM_CUDA(cudaMemcpyAsync(hi.data(), c->b_nbr.p, (size_t)total * 4, cudaMemcpyDeviceToHost, c->stream));
  TM_CUDA(cudaStreamSynchronize(c->stream));
  c->h_idx.resize((size_t)total);
  for (int64_t t = 0; t < total; t++) c->h_idx[t] = hi[t];
  return TM_OK;
}

// ================================================================ lattice path: windowed binning
// Periodic systems given as cell + lattice (tm_eval_lattice*, the slab phases).  The reference tessellates every atom
// into (2 ntess + 1)^3 images (Periodic.py:131-168: 648,000 positions for the 24,000-atom box) and lets the neighbour
// search sort out which of them matter.  Here only what can matter is ever touched: each real atom enumerates the
// lattice shifts (i, j, k) that put it inside the fractional window [wlo, whi] per axis -- the cell plus the interaction
// halo, or a slab rank's share of it -- and bins those images, carrying the reference's slot id  b * nreal + a  and the
// reference's left-to-right float64 arithmetic for the image position, so that everything downstream (neighbour rows,
// descriptor / pair / force kernels, the reference force convention on slot ids) is unchanged and bit-identical.
//   k_lat_count   counts per cell (all binned atoms) and per (element, cell) (centres), copies the real block
//   k_lat_scan    exclusive scans of [cell counts | centre counts e = 0 | e = 1 | ...]: chunk-local + chunk offsets, one launch
//   k_lat_scatter 32-byte records into their cells (arrival order)
//   k_lat_sort_rows  one warp per cell: order by slot id, assign the centre rows (element, cell order, slot)
struct LatBin {
  double L[9], ginv[9], wlo[3], whi[3], inv_n;
  int ntess;
  int slab_rank, slab_world;
  int64_t nreal;
  GridParams g;
};

// table segments are padded to 4 ints with room for the total behind the n entries
__device__ __forceinline__ int pad4(int n) { return (n + 4) & ~3; }
__device__ __forceinline__ double lat_frac(double x, double y, double z, const double* __restrict__ gi, int d) {
  return __fma_rn(z, gi[6 + d], __fma_rn(y, gi[3 + d], __dmul_rn(x, gi[d])));
}
// owner of a real atom among the slab ranks (floor(frac0 * world), clamped): same expression in every kernel
__device__ __forceinline__ int lat_owner(double f0, int world) {
  int o = (int)floor(f0 * (double)world);
  return o < 0 ? 0 : (o >= world ? world - 1 : o);
}
__device__ __forceinline__ int lat_block(int i, int j, int k, int nt) {
  const int side = 2 * nt + 1;
  const int lin = ((i + nt) * side + (j + nt)) * side + (k + nt), centre = (nt * side + nt) * side + nt;
  return lin < centre ? lin + 1 : (lin == centre ? 0 : lin);
}

// ONE enumeration shared by the counting and the scattering pass: calls fn(b, x, y, z, cid) for every binned image of
// real atom a (b = 0 is the atom itself).  LAT_SUB threads share an atom and take its images round-robin (an atom near a
// corner of the cell has 8 images inside the window, the average one 3: the atomics of the passes then run side by side).
#define LAT_SUB 4
template <typename F>
__device__ __forceinline__ void lat_enumerate(const LatBin& B, double x0, double y0, double z0, const double* f, int sub, F&& fn) {
  int lo[3], n[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    lo[d] = max(-B.ntess, (int)ceil(B.wlo[d] - f[d]));
    n[d] = max(0, min(B.ntess, (int)floor(B.whi[d] - f[d])) - lo[d] + 1);
  }
  const GridParams& g = B.g;
  const int nimg = n[0] * n[1] * n[2];
  for (int m = sub; m < nimg; m += LAT_SUB) {
    const int k = lo[2] + m % n[2], j = lo[1] + (m / n[2]) % n[1], i = lo[0] + m / (n[2] * n[1]);
    double x = x0, y = y0, z = z0;
    if (i | j | k) {   // coords_ + i*L0 + j*L1 + k*L2, left to right with separate roundings (numpy)
      double di = (double)i, dj = (double)j, dk = (double)k;
      x = __dadd_rn(__dadd_rn(__dadd_rn(x0, __dmul_rn(di, B.L[0])), __dmul_rn(dj, B.L[3])), __dmul_rn(dk, B.L[6]));
      y = __dadd_rn(__dadd_rn(__dadd_rn(y0, __dmul_rn(di, B.L[1])), __dmul_rn(dj, B.L[4])), __dmul_rn(dk, B.L[7]));
      z = __dadd_rn(__dadd_rn(__dadd_rn(z0, __dmul_rn(di, B.L[2])), __dmul_rn(dj, B.L[5])), __dmul_rn(dk, B.L[8]));
    }
    int cx = cell_coord(x, g.ox, g.inv_cell, g.gx), cy = cell_coord(y, g.oy, g.inv_cell, g.gy), cz = cell_coord(z, g.oz, g.inv_zcell, g.gz);
    fn(lat_block(i, j, k, B.ntess), x, y, z, (cx * g.gy + cy) * g.gz + cz);
  }
}

__device__ __forceinline__ int ele_index(const DevParams& P, int z) {
  int ei = -1;
  for (int k = 0; k < P.n_ele; k++) if (P.eles[k] == z) ei = k;
  return ei;
}

// cnt_all = [ncells (fine) cell counts | n_ele x ngroups centre counts per group of zdiv z bins], zeroed by the launcher
__global__ void k_lat_count(const double* __restrict__ xyz, const int32_t* __restrict__ Z, const __grid_constant__ LatBin B,
                            const __grid_constant__ DevParams P, double* __restrict__ pos, double* __restrict__ pos0, int32_t* __restrict__ Zo,
                            double* __restrict__ inv_n,
                            GridParams* __restrict__ gp, int32_t* __restrict__ cnt_all, int32_t* __restrict__ flags, int32_t* __restrict__ rowslot,
                            int64_t nrows, int32_t* __restrict__ rowofslot) {
  TM_PDL_PROLOGUE;
  const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  if (t0 == 0) { inv_n[0] = B.inv_n; *gp = B.g; }
  for (int64_t t = t0; t < nrows; t += stride) rowslot[t] = -1;
  const int ncells = B.g.ncells, zdiv = B.g.zdiv, ngroups = ncells / zdiv;
  for (int64_t t = t0; t < B.nreal * LAT_SUB; t += stride) {
    const int64_t a = t / LAT_SUB;
    const int sub = (int)(t - a * LAT_SUB);
    double x = xyz[3 * a], y = xyz[3 * a + 1], z = xyz[3 * a + 2];
    int zz = Z[a];
    if (sub == 0) {
      pos[3 * a] = x; pos[3 * a + 1] = y; pos[3 * a + 2] = z;   // the real block: read by k_charges (dipole)
      if (pos0) { pos0[3 * a] = x; pos0[3 * a + 1] = y; pos0[3 * a + 2] = z; }   // Verlet skin: where the lists were built
      Zo[a] = zz;
      rowofslot[a] = -1;
    }
    if (zz <= 0) continue;
    double f[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
      f[d] = lat_frac(x, y, z, B.ginv, d);
      if (sub == 0 && (f[d] < -1e-6 || f[d] > 1.0 + 1e-6)) atomicOr(flags, 8);   // not wrapped into the cell: the caller falls back
    }
    const int e = ele_index(P, zz);
    const bool centre = e >= 0 && (B.slab_world <= 1 || lat_owner(f[0], B.slab_world) == B.slab_rank);
    lat_enumerate(B, x, y, z, f, sub, [&](int b, double, double, double, int cid) {
      atomicAdd(&cnt_all[cid], 1);
      if (b == 0 && centre) atomicAdd(&cnt_all[pad4(ncells) + e * pad4(ngroups) + cid / zdiv], 1);
    });
  }
}

// Exclusive scans of the segments [n0 | n1 | n1 | ...] of `in` (segment s starts at s ? p0 + (s-1) p1 : 0, strides padded
// to 4 ints with at least one zero behind the n entries, so the scan value at index n is the segment total).  One launch:
// block b scans one 2048-element chunk locally and publishes the chunk total; the block that finishes last scans the
// chunk totals of every segment.  Readers combine the two levels: off(i) = local[i] + coff[chunk of i] (LatScan::at).
#define LSCAN_CH 2048
struct LatScan {
  const int32_t* local;   // chunk-local exclusive scans, same layout as `in`
  const int32_t* coff;    // exclusive scan of the chunk totals, per segment: [nch0 | nch1 | nch1 | ...]
  int p0, p1, nch0, nch1;
  __device__ __forceinline__ int at0(int i) const { return local[i] + coff[i / LSCAN_CH]; }                          // segment 0
  __device__ __forceinline__ int at(int e, int i) const { return local[p0 + e * p1 + i] + coff[nch0 + e * nch1 + i / LSCAN_CH]; }   // segment 1 + e
};
__global__ void __launch_bounds__(256) k_lat_scan(const int32_t* __restrict__ in, int32_t* __restrict__ local, int32_t* __restrict__ ctot,
                                                  int32_t* __restrict__ coff, int32_t* __restrict__ done, int n0, int n1, int nseg) {
  TM_PDL_PROLOGUE;
  __shared__ int32_t wsum[8];
  __shared__ int is_last;
  const int p0 = pad4(n0), p1 = pad4(n1);
  const int nch0 = (n0 + 1 + LSCAN_CH - 1) / LSCAN_CH, nch1 = (n1 + 1 + LSCAN_CH - 1) / LSCAN_CH;
  int blk = blockIdx.x, sgm = 0, ch = blk;
  if (blk >= nch0) { sgm = 1 + (blk - nch0) / nch1; ch = (blk - nch0) % nch1; }
  const int n = (sgm == 0 ? n0 : n1) + 1;                         // the zero behind the entries is scanned too
  const int64_t off = sgm == 0 ? 0 : (int64_t)p0 + (int64_t)(sgm - 1) * p1;
  const int base = ch * LSCAN_CH + threadIdx.x * 8;
  int32_t v[8], sum = 0;
  if (base + 7 < n) {
    const int4 a = *reinterpret_cast<const int4*>(in + off + base), b2 = *reinterpret_cast<const int4*>(in + off + base + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b2.x; v[5] = b2.y; v[6] = b2.z; v[7] = b2.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = (base + i < n) ? in[off + base + i] : 0;
  }
#pragma unroll
  for (int i = 0; i < 8; i++) sum += v[i];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int32_t t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  int32_t woff = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { if (i < w) woff += wsum[i]; tot += wsum[i]; }
  int32_t run = woff + inc - sum;
#pragma unroll
  for (int i = 0; i < 8; i++) { if (base + i < n) local[off + base + i] = run; run += v[i]; }
  if (threadIdx.x == 0) {
    ctot[blk] = tot;
    __threadfence();
    is_last = (atomicAdd(done, 1) == (int)gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // the last block: exclusive scan of the chunk totals of every segment (a few dozen values each), one warp per segment
  for (int sg = w; sg < nseg; sg += 8) {
    const int c0 = sg == 0 ? 0 : nch0 + (sg - 1) * nch1, nc = sg == 0 ? nch0 : nch1;
    int32_t carry = 0;
    for (int b0 = 0; b0 < nc; b0 += 32) {
      int32_t x = (b0 + lane < nc) ? ((volatile int32_t*)ctot)[c0 + b0 + lane] : 0, xi = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int32_t t = __shfl_up_sync(FULL, xi, o); if (lane >= o) xi += t; }
      if (b0 + lane < nc) coff[c0 + b0 + lane] = carry + xi - x;
      carry += __shfl_sync(FULL, xi, 31);
    }
  }
  if (threadIdx.x == 0) *done = 0;     // ready for the next step without a memset
}

__global__ void k_lat_scatter(const double* __restrict__ xyz, const int32_t* __restrict__ Z, const __grid_constant__ LatBin B,
                              const __grid_constant__ DevParams P, int32_t* __restrict__ cnt_all, const LatScan SC,
                              SAtom* __restrict__ sat) {
  TM_PDL_PROLOGUE;
  const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = t0; t < B.nreal * LAT_SUB; t += stride) {
    const int64_t a = t / LAT_SUB;
    const int sub = (int)(t - a * LAT_SUB);
    int zz = Z[a];
    if (zz <= 0) continue;
    double x = xyz[3 * a], y = xyz[3 * a + 1], z = xyz[3 * a + 2];
    double f[3];
#pragma unroll
    for (int d = 0; d < 3; d++) f[d] = lat_frac(x, y, z, B.ginv, d);
    const int e = ele_index(P, zz);
    lat_enumerate(B, x, y, z, f, sub, [&](int b, double xi, double yi, double zi, int cid) {
      int p = atomicSub(&cnt_all[cid], 1) - 1;     // the counts run back to zero: arrival rank inside the cell
      SAtom r;
      r.x = xi; r.y = yi; r.z = zi;
      r.slot = (int32_t)((int64_t)b * B.nreal + a);
      r.e = e;
      sat[SC.at0(cid) + p] = r;
    });
  }
}

// One warp per group of zdiv z bins (one cell edge of a column: ~10 atoms in liquid water): order the records of every bin
// by slot id (removes the arrival nondeterminism), then give every centre of the group its row: rows are ordered by
// (element, bin, slot) and each element's range starts at a multiple of TM_ROW_TILE.
// SC = the two-level scans of k_lat_scan: SC.at0(cid) = first record of bin cid (ncells + 1 entries), SC.at(e, g) per element
// offc_e[g] = centres of element e in the groups before g (ngroups + 1 entries, total last).
__global__ void k_lat_sort_rows(const __grid_constant__ LatBin B, int n_ele, const LatScan SC, int32_t* __restrict__ cstart,
                                SAtom* __restrict__ sat, int32_t* __restrict__ rowmeta, int32_t* __restrict__ rowslot, int32_t* __restrict__ rowsidx,
                                int32_t* __restrict__ rowofslot, int64_t nrows, int32_t* __restrict__ flags) {
  TM_PDL_PROLOGUE;
  const int ncells = B.g.ncells, zdiv = B.g.zdiv, ngroups = ncells / zdiv;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  int ebase[TM_MAX_ELE];
  {
    int base = 0, total = 0;
#pragma unroll
    for (int e = 0; e < TM_MAX_ELE; e++) {
      int cnt = (e < n_ele) ? SC.at(e, ngroups) : 0;
      ebase[e] = base;
      if (blockIdx.x == 0 && threadIdx.x == 0) { rowmeta[2 * e] = base; rowmeta[2 * e + 1] = cnt; }
      base += ((cnt + TM_ROW_TILE - 1) / TM_ROW_TILE) * TM_ROW_TILE;
      total += cnt;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { rowmeta[2 * TM_MAX_ELE] = total; rowmeta[2 * TM_MAX_ELE + 1] = base; }
  }
  for (int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; grp <= ngroups; grp += warps) {
    if (grp == ngroups) { if (lane == 0) cstart[ncells] = SC.at0(ncells); continue; }
    const int c0 = grp * zdiv;
    // bin boundaries of the group (zdiv <= 31): lane k holds off_all[c0 + k]
    const int mybound = (lane <= zdiv) ? SC.at0(c0 + lane) : 0x7fffffff;
    if (lane < zdiv) cstart[c0 + lane] = mybound;
    const int b = __shfl_sync(FULL, mybound, 0), e = __shfl_sync(FULL, mybound, zdiv), n = e - b;
    if (n <= 0) continue;
    if (n > 32) {     // dense group: serial insertion sort per bin by lane 0, then the passes below only assign rows
      if (lane == 0) {
        for (int k = 0; k < zdiv; k++) {
          const int bb = SC.at0(c0 + k), be = SC.at0(c0 + k + 1);
          for (int i = bb + 1; i < be; i++) {
            SAtom v = sat[i];
            int j = i - 1;
            while (j >= bb && sat[j].slot > v.slot) { sat[j + 1] = sat[j]; j--; }
            sat[j + 1] = v;
          }
        }
      }
      __syncwarp();
    }
    int seen[TM_MAX_ELE];
#pragma unroll
    for (int q = 0; q < TM_MAX_ELE; q++) seen[q] = 0;
    for (int p0 = 0; p0 < n; p0 += 32) {
      const int cnt = min(32, n - p0);
      SAtom mine;
      mine.slot = 0x7fffffff; mine.e = -1; mine.x = mine.y = mine.z = 0.0;
      if (lane < cnt) mine = sat[b + p0 + lane];
      int rank = lane;
      if (n <= 32) {
        // sort key (bin, slot): the bin of a record is known from its position (the scatter pass filled bin by bin)
        int bin = 0;
        for (int k = 1; k < zdiv; k++) bin += (b + lane >= __shfl_sync(FULL, mybound, k)) ? 1 : 0;
        const long long key = (lane < cnt) ? (((long long)bin << 32) | (unsigned int)mine.slot) : 0x7fffffffffffffffll;
        rank = 0;
        for (int k = 0; k < cnt; k++) {
          long long other = __shfl_sync(FULL, key, k);
          rank += (other < key) ? 1 : 0;
        }
        __syncwarp();
        if (lane < cnt) sat[b + rank] = mine;
      }
      // centres: real slots owned by this rank
      bool centre = false;
      if (lane < cnt && mine.e >= 0 && mine.slot < B.nreal) {
        centre = true;
        if (B.slab_world > 1) centre = lat_owner(lat_frac(mine.x, mine.y, mine.z, B.ginv, 0), B.slab_world) == B.slab_rank;
      }
#pragma unroll
      for (int q = 0; q < TM_MAX_ELE; q++) {
        if (q >= n_ele) break;
        // centres of element q of this pass as a bit mask over SORTED positions: those below this record come first
        const unsigned mq = __reduce_or_sync(FULL, (centre && mine.e == q) ? (1u << rank) : 0u);
        if (centre && mine.e == q) {
          int row = ebase[q] + SC.at(q, grp) + seen[q] + __popc(mq & ((1u << rank) - 1u));
          if (row < nrows) {
            rowslot[row] = mine.slot;
            rowsidx[row] = b + p0 + rank;
            rowofslot[mine.slot] = row;
          } else {
            atomicOr(flags, 64);
          }
        }
        seen[q] += __popc(mq);
      }
    }
  }
}

int tm_launch_lattice_bin(tm_ctx* c, const SysView& s) {
  int rc;
  const int n_ele = c->hp.n_ele;
  const int64_t ncells = s.hgrid.ncells, ngroups = ncells / s.hgrid.zdiv;
  auto hpad4 = [](int64_t n) { return (n + 4) & ~(int64_t)3; };
  const int64_t nall = hpad4(ncells) + (int64_t)n_ele * hpad4(ngroups);
  if ((rc = tm_buf(c, c->b_pos, (size_t)s.nreal * 24))) return rc;
  if ((rc = tm_buf(c, c->b_Z, (size_t)s.nreal * 4))) return rc;
  if ((rc = tm_buf(c, c->b_natom, 8))) return rc;
  if (c->skin > 0.0 && (rc = tm_buf(c, c->b_pos0, (size_t)s.nreal * 24))) return rc;
  if ((rc = tm_buf(c, c->b_cntall, (size_t)(nall + 8) * 4))) return rc;
  if ((rc = tm_buf(c, c->b_offall, (size_t)(nall + n_ele + 8) * 4))) return rc;
  if ((rc = tm_buf(c, c->b_cstart, (size_t)(ncells + 8) * 4))) return rc;
  if ((rc = tm_buf(c, c->b_satom, (size_t)s.nslots * sizeof(SAtom)))) return rc;
  if ((rc = tm_buf(c, c->b_grid, sizeof(GridParams)))) return rc;
  if ((rc = tm_buf(c, c->b_rowmeta, (2 * TM_MAX_ELE + 2) * 4))) return rc;
  if ((rc = tm_buf(c, c->b_rowslot, (size_t)s.nrows * 4))) return rc;
  if ((rc = tm_buf(c, c->b_rowsidx, (size_t)s.nrows * 4))) return rc;
  if ((rc = tm_buf(c, c->b_rowofslot, (size_t)s.nreal * 4))) return rc;
  LatBin B;
  memcpy(B.L, s.lat.v, 72);
  memcpy(B.ginv, s.ginv, 72);
  for (int d = 0; d < 3; d++) { B.wlo[d] = s.wlo[d]; B.whi[d] = s.whi[d]; }
  B.inv_n = s.lat.v[9];
  B.ntess = s.lat_ntess;
  B.slab_rank = s.slab_rank; B.slab_world = s.slab_world;
  B.nreal = s.nreal;
  B.g = s.hgrid;
  int32_t* cnt_all = (int32_t*)c->b_cntall.p;
  int32_t* off_all = (int32_t*)c->b_offall.p;
  const int nch0 = (int)((ncells + 1 + LSCAN_CH - 1) / LSCAN_CH), nch1 = (int)((ngroups + 1 + LSCAN_CH - 1) / LSCAN_CH);
  const int nchunks = nch0 + n_ele * nch1;
  if ((rc = tm_buf(c, c->b_lscan, (size_t)(2 * nchunks + 16) * 4))) return rc;    // [done counter | chunk totals | chunk offsets]
  int32_t* done = (int32_t*)c->b_lscan.p;   // at a FIXED place (the chunk count changes with the grid): zero from the allocation, reset by the kernel itself
  int32_t* ctot = done + 4;
  int32_t* coff = ctot + nchunks;
  LatScan SC{off_all, coff, (int)hpad4(ncells), (int)hpad4(ngroups), nch0, nch1};
  TM_CUDA(cudaMemsetAsync(cnt_all, 0, (size_t)(nall + 8) * 4, c->stream));
  int blocks = (int)((std::max<int64_t>(s.nreal * LAT_SUB, s.nrows) + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  TM_LAUNCH(k_lat_count, blocks, 256, 0, c->stream, s.xyz_real, s.Z_real, B, c->hp, (double*)c->b_pos.p, c->skin > 0.0 ? (double*)c->b_pos0.p : (double*)nullptr,
            (int32_t*)c->b_Z.p, (double*)c->b_natom.p,
                                             (GridParams*)c->b_grid.p, cnt_all, (int32_t*)c->b_flags.p, (int32_t*)c->b_rowslot.p, s.nrows,
                                             (int32_t*)c->b_rowofslot.p);
  c->launches++;
  TM_LAUNCH(k_lat_scan, nchunks, 256, 0, c->stream, cnt_all, off_all, ctot, coff, done, (int)ncells, (int)ngroups, 1 + n_ele);
  c->launches++;
  int sblocks = (int)((s.nreal * LAT_SUB + 255) / 256);
  if (sblocks > 148 * 8) sblocks = 148 * 8;
  TM_LAUNCH(k_lat_scatter, sblocks, 256, 0, c->stream, s.xyz_real, s.Z_real, B, c->hp, cnt_all, SC, (SAtom*)c->b_satom.p);
  int cblocks = (int)(((ngroups + 1) * 32 + 255) / 256);
  if (cblocks > 148 * 16) cblocks = 148 * 16;
  TM_LAUNCH(k_lat_sort_rows, cblocks, 256, 0, c->stream, B, n_ele, SC, (int32_t*)c->b_cstart.p, (SAtom*)c->b_satom.p, (int32_t*)c->b_rowmeta.p,
                                                  (int32_t*)c->b_rowslot.p, (int32_t*)c->b_rowsidx.p, (int32_t*)c->b_rowofslot.p, s.nrows,
                                                  (int32_t*)c->b_flags.p);
  c->launches += 2;
  TM_CUDA(cudaGetLastError());
  return TM_OK;
}

// ---------------------------------------------------------------- Verlet-skin reuse
// New positions into the existing cell-sorted records (cells, neighbour rows and centre rows stay as built): record i
// belongs to slot b nreal + a, i.e. to real atom a shifted by the lattice vector of image block b, recomputed with the
// reference's arithmetic from the atom's new position.  Also refreshes the real block (dipole) and checks the contract:
// no atom further than skin / 2 from where it was when the rows were built (flag 128).
__global__ void k_lat_refresh(const double* __restrict__ xyz, const __grid_constant__ LatBin B, const int32_t* __restrict__ cstart,
                              SAtom* __restrict__ sat, double* __restrict__ pos, const double* __restrict__ pos0, double half_skin2,
                              int32_t* __restrict__ flags) {
  TM_PDL_PROLOGUE;
  const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t a = t0; a < B.nreal; a += stride) {
    double x = xyz[3 * a], y = xyz[3 * a + 1], z = xyz[3 * a + 2];
    pos[3 * a] = x; pos[3 * a + 1] = y; pos[3 * a + 2] = z;
    double dx = x - pos0[3 * a], dy = y - pos0[3 * a + 1], dz = z - pos0[3 * a + 2];
    if (dx * dx + dy * dy + dz * dz > half_skin2) atomicOr(flags + 8, 128);   // sticky word: survives the per-step reset
  }
  const int ntot = cstart[B.g.ncells];
  const int nt = B.ntess, side = 2 * nt + 1, centre = (nt * side + nt) * side + nt;
  for (int64_t r = t0; r < ntot; r += stride) {
    const int slot = sat[r].slot;
    const int b = (int)(slot / B.nreal);
    const int64_t a = slot - (int64_t)b * B.nreal;
    double x = xyz[3 * a], y = xyz[3 * a + 1], z = xyz[3 * a + 2];
    if (b > 0) {
      const int lin = (b - 1 < centre) ? (b - 1) : b;
      const double dk = (double)(lin % side - nt), dj = (double)((lin / side) % side - nt), di = (double)(lin / (side * side) - nt);
      const double x0 = x, y0 = y, z0 = z;
      x = __dadd_rn(__dadd_rn(__dadd_rn(x0, __dmul_rn(di, B.L[0])), __dmul_rn(dj, B.L[3])), __dmul_rn(dk, B.L[6]));
      y = __dadd_rn(__dadd_rn(__dadd_rn(y0, __dmul_rn(di, B.L[1])), __dmul_rn(dj, B.L[4])), __dmul_rn(dk, B.L[7]));
      z = __dadd_rn(__dadd_rn(__dadd_rn(z0, __dmul_rn(di, B.L[2])), __dmul_rn(dj, B.L[5])), __dmul_rn(dk, B.L[8]));
    }
    sat[r].x = x; sat[r].y = y; sat[r].z = z;
  }
}

int tm_launch_lattice_refresh(tm_ctx* c, const SysView& s) {
  LatBin B;
  memcpy(B.L, s.lat.v, 72);
  memcpy(B.ginv, s.ginv, 72);
  for (int d = 0; d < 3; d++) { B.wlo[d] = s.wlo[d]; B.whi[d] = s.whi[d]; }
  B.inv_n = s.lat.v[9];
  B.ntess = s.lat_ntess;
  B.slab_rank = s.slab_rank; B.slab_world = s.slab_world;
  B.nreal = s.nreal;
  B.g = s.hgrid;
  if (!c->b_pos0.p || !c->b_satom.p) { tm_set_error("no neighbour rows to reuse"); return TM_ESTATE; }
  int blocks = (int)((4 * s.nreal + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  const double hs = 0.5 * c->skin;
  TM_LAUNCH(k_lat_refresh, blocks, 256, 0, c->stream, s.xyz_real, B, (const int32_t*)c->b_cstart.p, (SAtom*)c->b_satom.p, (double*)c->b_pos.p,
            (const double*)c->b_pos0.p, hs * hs, (int32_t*)c->b_flags.p);
  c->launches++;
  TM_CUDA(cudaGetLastError());
  return TM_OK;
}
