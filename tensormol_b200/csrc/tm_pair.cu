// K5: charge neutralisation + damped-shifted-force / ELU Coulomb + Grimme-C6 vdW pair kernel.
//
// Restates on the device:
//   charge neutralisation + dipole        TFMolInstanceDirect.py:5274-5279 (periodic :5881-5893)
//   TFCoulombEluSRDSFLR                    RawSymFunc.py:1307-1359
//   TFVdwPolyLR / TFVdwPolyLRWithEle       RawSymFunc.py:1361-1465 (double Bohr scaling kept, Q6)
// and the part of tf.gradients that flows through them.  The 15 A pair list (34 M rows at 24k atoms,
// Neighbors.py:323-342) is never materialised: every centre walks the cell grid.
//
// Energy convention: E = 1/2 sum over ORDERED pairs (i centre, j any slot) -- equal to the reference's
// i<j sum for aperiodic input (TFMolManage.py:1311-1312) and to its "/2" periodic form
// (TFMolInstanceDirect.py:5896, 5824).  Gradient on a real row a: sum_j w_j d e_aj/d x_a with w_j = 1
// for real j and 1/2 for image j (image rows are independent variables whose gradient the reference
// discards).  dEcc/dq_a = sum_j q_j kappa_aj uses the image symmetry of a full
// tessellation (Periodic.py:131-168), which every reference flow provides (Q12).
//
// Mapping: one warp per centre; lanes stride over contiguous z-runs of the cell-sorted copy, candidates
// inside the cutoff are compacted (ballot/popc) into a per-warp shared queue and evaluated 32 at a time
// so the erfc/exp work runs on full warps.
#include "tm_internal.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <utility>
#include <vector>

#define FULL 0xffffffffu
#define PAIR_WARPS 8
#define QCAP 96

// ---- charges -----------------------------------------------------------------------------------
// y_charge[row] -> qraw_slot[slot]; per-molecule sums in double
__global__ void k_qraw_scatter(const float* __restrict__ y, const int32_t* __restrict__ rowslot, int64_t nrows, double* __restrict__ qraw_slot) {
  TM_PDL_PROLOGUE;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x) {
    int s = rowslot[r];
    if (s >= 0) qraw_slot[s] = (double)y[r];
  }
}

// molsum[m] = sum over real slots of molecule m (one block per molecule chunk)
__global__ void k_mol_sum(const double* __restrict__ v, const int32_t* __restrict__ Z, int64_t maxnatom, int64_t nvalid_per_mol, double* __restrict__ molacc, int field) {
  TM_PDL_PROLOGUE;
  int m = blockIdx.x;
  double s = 0.0;
  for (int64_t a = blockIdx.y * (int64_t)blockDim.x + threadIdx.x; a < nvalid_per_mol; a += (int64_t)gridDim.y * blockDim.x) {
    int64_t slot = (int64_t)m * maxnatom + a;
    if (Z[slot] > 0) s += v[slot];
  }
  __shared__ double sh[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if (threadIdx.x == 0) atomicAdd(&molacc[16 * m + field], s);   // molacc is zeroed at the start of the evaluation
  }
}

// One launch for the two consumers of the neutralised charges; block (x, m) works on molecule m.
//  (1) q_slot = qraw - molsum * inv_n[m]  for the first nq slots of each molecule (padding included, Q11);
//      dipole[m] += q * B * x   (TFMolInstanceDirect.py:5278-5279)
//  (2) the candidate records of the pair kernel, one float4 per cell-sorted atom of the molecule: position relative to
//      the grid origin in fp32 (only used for the in/out-of-cutoff test, where the kernel vanishes) and the charge
//      (images inherit the charge of slot % nreal, TFMolInstanceDirect.py:5892-5893).
__global__ void k_charges(const double* __restrict__ qraw_slot, double* __restrict__ molacc, const double* __restrict__ inv_n,
                          const double* __restrict__ pos, const int32_t* __restrict__ Z, int64_t maxnatom, int64_t nq_per_mol,
                          double* __restrict__ q_slot, const SAtom* __restrict__ sat, const int32_t* __restrict__ cstart,
                          const GridParams* __restrict__ gp, int64_t nreal, int periodic, float4* __restrict__ pq, uint8_t* __restrict__ pe,
                          int self_sum) {
  TM_PDL_PROLOGUE;
  int m = blockIdx.y;
  // self_sum (one molecule, slab runs): every block forms sum q_raw by itself, all in the same order, instead of a separate
  // reduction launch; block 0 publishes it
  __shared__ double s_part[32];
  __shared__ double s_total;
  if (self_sum) {
    double t = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;     // four loads in flight per thread (the loop is pure L2 latency)
    const int64_t bd = blockDim.x;
    for (int64_t a = threadIdx.x; a < nq_per_mol; a += 4 * bd) {
      t += (Z[a] > 0) ? qraw_slot[a] : 0.0;
      if (a + bd < nq_per_mol) t1 += (Z[a + bd] > 0) ? qraw_slot[a + bd] : 0.0;
      if (a + 2 * bd < nq_per_mol) t2 += (Z[a + 2 * bd] > 0) ? qraw_slot[a + 2 * bd] : 0.0;
      if (a + 3 * bd < nq_per_mol) t3 += (Z[a + 3 * bd] > 0) ? qraw_slot[a + 3 * bd] : 0.0;
    }
    t = (t + t1) + (t2 + t3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tt = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) tt += s_part[w];
      s_total = tt;
      if (blockIdx.x == 0) molacc[4] = tt;
    }
    __syncthreads();
  }
  double mean = (self_sum ? s_total : molacc[16 * m + 4]) * inv_n[m];
  double d0 = 0, d1 = 0, d2 = 0;
  for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < nq_per_mol; a += (int64_t)gridDim.x * blockDim.x) {
    int64_t slot = (int64_t)m * maxnatom + a;
    double q = ((Z[slot] > 0) ? qraw_slot[slot] : 0.0) - mean;
    q_slot[slot] = q;
    d0 += q * TM_BOHRPERA * pos[3 * slot];
    d1 += q * TM_BOHRPERA * pos[3 * slot + 1];
    d2 += q * TM_BOHRPERA * pos[3 * slot + 2];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    d0 += __shfl_xor_sync(FULL, d0, o);
    d1 += __shfl_xor_sync(FULL, d1, o);
    d2 += __shfl_xor_sync(FULL, d2, o);
  }
  if ((threadIdx.x & 31) == 0 && (d0 != 0.0 || d1 != 0.0 || d2 != 0.0)) {
    atomicAdd(&molacc[16 * m + 6], d0);
    atomicAdd(&molacc[16 * m + 7], d1);
    atomicAdd(&molacc[16 * m + 8], d2);
  }
  // (2): every binned atom has Z > 0, so its charge is qraw - mean (the value (1) stores for its slot)
  GridParams g = *gp;
  int ib = cstart[m * g.ncell_mol], ie = cstart[(m + 1) * g.ncell_mol];
  // positions relative to the CENTRE of the grid (halves the magnitude, i.e. the fp32 rounding step: 3.8e-6 A for a 92 A grid)
  const double mx = g.ox + 0.5 * g.gx * g.cell, my = g.oy + 0.5 * g.gy * g.cell, mz = g.oz + 0.5 * g.gz * g.zcell;
  for (int i = ib + blockIdx.x * blockDim.x + threadIdx.x; i < ie; i += gridDim.x * blockDim.x) {
    SAtom a = sat[i];
    int64_t s = a.slot;
    bool img = false;
    if (periodic) { img = s >= nreal; s = s % nreal; }
    pq[i] = make_float4((float)(a.x - mx), (float)(a.y - my), (float)(a.z - mz), (float)(qraw_slot[s] - mean));
    pe[i] = (uint8_t)((a.e & 7) | (img ? 0x80 : 0));
  }
}

int tm_launch_charges(tm_ctx* c, const SysView& s) {
  int rc;
  int64_t nq = s.periodic ? s.nreal : s.nslots;           // slots that carry an own charge
  int64_t nq_per_mol = s.periodic ? s.nreal : s.maxnatom;
  if ((rc = tm_buf(c, c->b_q, (size_t)nq * 8 * 2))) return rc;  // [qraw_slot | q_slot]
  if ((rc = tm_buf(c, c->b_qs, (size_t)s.nslots * 16))) return rc;
  if ((rc = tm_buf(c, c->b_pe, (size_t)s.nslots))) return rc;
  double* qraw = (double*)c->b_q.p;
  double* q = qraw + nq;
  // molacc (zeroed by the caller at the start of the evaluation), stride 16 doubles per molecule:
  //   [0] Etotal [1] Ebp [2] Ecc [3] Evdw [4] sum q_raw [5] sum dE/dq [6..8] dipole
  double* molacc = (double*)c->b_molacc.p;
  // slab mode: qraw_slot was already combined over the ranks and written into b_q; otherwise the tensor-core forward
  // pass has scattered q_raw and summed it per molecule (k_y_reduce), and the fp32 mode does both here
  const bool fused = c->y_fused && !s.slab_api;
  const int self_sum = 0;   // (every block summing q_raw by itself measured slower than the separate k_mol_sum launch: 28 us vs 15)
  if (!fused && !self_sum) {
    if (!s.slab_api) {
      int blocks = (int)((s.nrows + 255) / 256);
      TM_CUDA(cudaMemsetAsync(qraw, 0, (size_t)nq * 8, c->stream));
      TM_LAUNCH(k_qraw_scatter, blocks, 256, 0, c->stream, (const float*)c->b_y[TM_NET_CHARGE].p, (const int32_t*)c->b_rowslot.p, s.nrows, qraw);
      c->launches++;
    }
    dim3 gms((unsigned)s.nmol, (unsigned)std::max<int64_t>(1, std::min<int64_t>((nq_per_mol + 2047) / 2048, 64)));
    TM_LAUNCH(k_mol_sum, gms, 256, 0, c->stream, qraw, (const int32_t*)c->b_Z.p, s.maxnatom, nq_per_mol, molacc, 4);
    c->launches++;
  }
  int64_t per_mol = std::max<int64_t>(nq_per_mol, s.nslots / std::max<int64_t>(1, s.nmol));
  // the lattice path bins a few images per atom at most (window = cell + halo), and a block that sums q_raw by itself
  // should not be one of thousands
  int64_t cap_blocks = self_sum ? 74 : (s.lat_bin ? 148 * 2 : 148 * 8);
  dim3 g((unsigned)std::max<int64_t>(1, std::min<int64_t>((per_mol + 255) / 256, cap_blocks)), (unsigned)s.nmol);
  TM_LAUNCH(k_charges, g, self_sum ? 1024 : 256, 0, c->stream, qraw, molacc, (const double*)c->b_natom.p, (const double*)c->b_pos.p, (const int32_t*)c->b_Z.p, s.maxnatom,
                                      nq_per_mol, q, (const SAtom*)c->b_satom.p, (const int32_t*)c->b_cstart.p, (const GridParams*)c->b_grid.p, s.nreal,
                                      s.periodic, (float4*)c->b_qs.p, (uint8_t*)c->b_pe.p, self_sum);
  c->launches++;
  TM_CUDA(cudaGetLastError());
  return TM_OK;
}

// ---- pair kernel -------------------------------------------------------------------------------
struct PairAcc {
  float ecc, evdw, dedq, gx, gy, gz;
};

// erfc(x) = exp(-x^2) * erfcx(x); erfcx is smooth on the LR branch's range [alpha R_sr, alpha R_lr] and is evaluated as a
// degree-11 polynomial in u in [-1,1] fitted on the host (tm_api.cu, max relative error < 1e-7), sharing the exp(-x^2)
// that the derivative needs anyway.  All unit factors (Bohr/Angstrom, log2 e, 1/B in erfc(aR)/R, the double Bohr scaling
// of the vdW term) are folded into the DevParams pk_* constants; MUFU: rsqrt, ex2, rcp.
//   d2, (dx,dy,dz) = x_j - x_i in Angstrom; c6r / rs12r = this centre's rows of pk_c6 / pk_rs12 (shared memory);
//   accumulates e_ij pieces and  d e_ij / d x_i = -(dE/dr) (x_j - x_i)/r  weighted by wj.
__device__ __forceinline__ float ex2f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
template <bool ECC, bool VDW>
__device__ __forceinline__ void pair_eval(const DevParams& P, float d2, float dx, float dy, float dz, float qi, float qj, float c6, float rs12,
                                          float wj, PairAcc& A) {
  float ir = rsqrt_approx(d2);    // 1/r (2 ulp: d2 is never denormal for distinct atoms)
  float r = d2 * ir;
  float de = 0.f;                 // dE/dr (Hartree / Angstrom)
  if (ECC) {
    float kap = 0.f, bdk = 0.f;   // kappa and B dkappa/dR
    if (d2 > P.pk_rsr2) {
      if (d2 <= P.pk_rlr2) {
        float ex = ex2f(P.pk_cex * d2);
        float u = fmaf(r, P.pk_ua, P.pk_ub);
        float pz = P.pk_pc[11];
#pragma unroll
        for (int k = 10; k >= 0; k--) pz = fmaf(pz, u, P.pk_pc[k]);
        float eir = pz * ex * ir;                       // erfc(aR)/R
        kap = eir + fmaf(r, P.pk_ka, P.pk_kb);
        bdk = fmaf(-ir, fmaf(P.pk_c2, ex, eir), P.pk_BZY);
      }
    } else {
      float ex = ex2f(fmaf(r, P.pk_ea, P.pk_eb));
      kap = fmaf(P.elu_a, ex, P.pk_ec);
      bdk = P.pk_belu * ex;
    }
    float qq = qi * qj;
    A.ecc = fmaf(qq, kap, A.ecc);
    A.dedq = fmaf(qj, kap, A.dedq);
    de = qq * bdk;
  }
  if (VDW) {
    float id2 = ir * ir;
    float id6 = id2 * id2 * id2;
    float X = rs12 * id6 * id6;                         // 6 x^-12
    float damp = rcp_approx(1.0f + X);
    float fd = c6 * id6 * damp;                         // C6 / R'^6 * damp
    float t = fminf(r * P.pk_ta, 1.0f);                 // switch argument, clamped: S(1) = 1, S'(1) = 0
    float S = t * t * fmaf(-2.0f, t, 3.0f);
    float bds = 6.0f * t * (1.0f - t) * P.pk_ta;        // B^2 dS/dR'
    A.evdw = fmaf(-S, fd, A.evdw);
    de -= fd * fmaf(S * ir, fmaf(12.0f * X, damp, -6.0f), bds);
  }
  float sc = -wj * de * ir;
  A.gx = fmaf(sc, dx, A.gx); A.gy = fmaf(sc, dy, A.gy); A.gz = fmaf(sc, dz, A.gz);
}

// One warp per centre.  The cell columns (x, y) that can hold a partner are resolved 32 at a time, one column per lane
// (its z-range from the centre's actual position: contiguous run [b, e) of the cell-sorted copy), then the warp walks
// the runs; candidates inside the cutoff are compacted into the shared queue and evaluated 32 at a time.
template <bool ECC, bool VDW>
__global__ void __launch_bounds__(PAIR_WARPS * 32)
k_pair(const SAtom* __restrict__ sat, const float4* __restrict__ pq, const int32_t* __restrict__ cstart, const GridParams* __restrict__ gp,
       const int32_t* __restrict__ rowsidx, const int32_t* __restrict__ rowslot, int64_t nrows, int64_t maxnatom, int nreal_slots,
       const __grid_constant__ DevParams P, int do_force, float cutoff_A, int split, double* __restrict__ dedq_slot, float* __restrict__ F,
       double* __restrict__ molacc) {
  TM_PDL_PROLOGUE;
  __shared__ int q_j[PAIR_WARPS][QCAP];
  __shared__ float s_c6[PAIR_WARPS][TM_MAX_ELE], s_rs12[PAIR_WARPS][TM_MAX_ELE];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // `split` warps share one centre (few centres: slab runs, small systems); warp `sub` takes every split-th column
  int64_t gw = (int64_t)blockIdx.x * PAIR_WARPS + warp;
  int64_t row = gw / split;
  int sub = (int)(gw - row * split);
  if (row >= nrows) return;
  int slot = rowslot[row];
  if (slot < 0) return;
  GridParams g = *gp;
  int si = rowsidx[row];
  SAtom ci = sat[si];
  float4 pi = pq[si];
  float qi = pi.w;
  int ei = ci.e;
  if (lane < TM_MAX_ELE) { s_c6[warp][lane] = P.pk_c6[ei][lane]; s_rs12[warp][lane] = P.pk_rs12[ei][lane]; }
  __syncwarp();
  int m = (int)(slot / maxnatom);
  float cell = (float)g.cell, icell = (float)g.inv_cell, izcell = (float)g.inv_zcell;
  float rc2 = cutoff_A * cutoff_A;
  const unsigned lt_mask = (1u << lane) - 1u;
  // the candidate records are relative to the grid centre; the column arithmetic wants the origin
  const float ox_ = (float)(ci.x - g.ox), oy_ = (float)(ci.y - g.oy), oz_ = (float)(ci.z - g.oz);
  // cell columns that can hold a partner, from the centre's actual position (not its cell): |dx| <= rc (+ the Verlet skin:
  // with reused lists an atom may have left the cell it was binned in by up to skin / 2)
  const float rcw = cutoff_A + P.skin, rcw2 = rcw * rcw;
  int x0 = max(0, (int)floorf((ox_ - rcw) * icell)), x1 = min(g.gx - 1, (int)floorf((ox_ + rcw) * icell));
  int y0 = max(0, (int)floorf((oy_ - rcw) * icell)), y1 = min(g.gy - 1, (int)floorf((oy_ + rcw) * icell));
  int ny = y1 - y0 + 1, ncol = (x1 - x0 + 1) * ny;
  PairAcc A = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int qn = 0;
  // image partners: 1/2 under the reference's convention (header); 1 when the image rows are folded onto their real
  // atom (TM_F_FOLD_IMAGES, do_force bit 1): then the row gradient is the derivative of the periodic energy
  const float w_img = (do_force & 2) ? 1.0f : 0.5f;
  auto eval = [&](int j) {
    SAtom a = sat[j];
    float ddx = (float)(a.x - ci.x), ddy = (float)(a.y - ci.y), ddz = (float)(a.z - ci.z);
    float d2 = ddx * ddx + ddy * ddy + ddz * ddz;
    pair_eval<ECC, VDW>(P, d2, ddx, ddy, ddz, qi, pq[j].w, s_c6[warp][a.e], s_rs12[warp][a.e], (a.slot < nreal_slots) ? 1.0f : w_img, A);
  };
  for (int c0 = 0; c0 * split < ncol; c0 += 32) {
    // lane -> column (c0+lane)*split + sub: run [cb, ce) or empty
    int cb = 0, ce = 0;
    int cidx = (c0 + lane) * split + sub;
    if (cidx < ncol) {
      int x = x0 + cidx / ny, y = y0 + cidx % ny;
      float lx = fmaxf(0.f, fmaxf(x * cell - ox_, ox_ - (x + 1) * cell));   // distance from the centre to the column's slab in x
      float ly = fmaxf(0.f, fmaxf(y * cell - oy_, oy_ - (y + 1) * cell));
      float rem = rcw2 - lx * lx - ly * ly;
      if (rem > 0.f) {
        float zr = sqrtf(rem);
        int z0 = max(0, (int)floorf((oz_ - zr) * izcell)), z1 = min(g.gz - 1, (int)floorf((oz_ + zr) * izcell));
        if (z1 >= z0) {
          int cbase = m * g.ncell_mol + (x * g.gy + y) * g.gz;
          cb = cstart[cbase + z0];
          ce = cstart[cbase + z1 + 1];
        }
      }
    }
    unsigned live = __ballot_sync(FULL, ce > cb);
    while (live) {
      int k = __ffs(live) - 1;
      live &= live - 1;
      int b = __shfl_sync(FULL, cb, k), e = __shfl_sync(FULL, ce, k);
      int j0 = b;
      // 64 candidates per pass while the run is long enough (two loads in flight, one set of loop / queue bookkeeping) ...
      for (; j0 + 32 < e; j0 += 64) {
        int ja = j0 + lane, jb = ja + 32;
        float4 pa = pq[ja];
        float d2b = 3.0e38f;
        if (jb < e) {
          float4 pb = pq[jb];
          float bx = pb.x - pi.x, by = pb.y - pi.y, bz = pb.z - pi.z;
          d2b = bx * bx + by * by + bz * bz;
        }
        float ddx = pa.x - pi.x, ddy = pa.y - pi.y, ddz = pa.z - pi.z;
        float d2a = ddx * ddx + ddy * ddy + ddz * ddz;
        bool oka = (d2a < rc2) && (ja != si), okb = (d2b < rc2) && (jb != si);
        unsigned mka = __ballot_sync(FULL, oka), mkb = __ballot_sync(FULL, okb);
        int na = __popc(mka);
        if (oka) q_j[warp][qn + __popc(mka & lt_mask)] = ja;
        if (okb) q_j[warp][qn + na + __popc(mkb & lt_mask)] = jb;
        qn += na + __popc(mkb);
        __syncwarp();
        while (qn >= 32) {
          eval(q_j[warp][qn - 32 + lane]);
          qn -= 32;
          __syncwarp();
        }
      }
      // ... and a last pass of up to 32
      if (j0 < e) {
        int j = j0 + lane;
        float d2 = 3.0e38f;
        if (j < e) {
          float4 pj = pq[j];
          float ddx = pj.x - pi.x, ddy = pj.y - pi.y, ddz = pj.z - pi.z;
          d2 = ddx * ddx + ddy * ddy + ddz * ddz;
        }
        bool ok = (d2 < rc2) && (j != si);
        unsigned mk = __ballot_sync(FULL, ok);
        if (ok) q_j[warp][qn + __popc(mk & lt_mask)] = j;
        qn += __popc(mk);
        __syncwarp();
        if (qn >= 32) {
          eval(q_j[warp][qn - 32 + lane]);
          qn -= 32;
          __syncwarp();
        }
      }
    }
  }
  if (lane < qn) eval(q_j[warp][lane]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    A.ecc += __shfl_xor_sync(FULL, A.ecc, o);
    A.evdw += __shfl_xor_sync(FULL, A.evdw, o);
    A.dedq += __shfl_xor_sync(FULL, A.dedq, o);
    A.gx += __shfl_xor_sync(FULL, A.gx, o);
    A.gy += __shfl_xor_sync(FULL, A.gy, o);
    A.gz += __shfl_xor_sync(FULL, A.gz, o);
  }
  if (lane == 0) {
    if (split == 1) dedq_slot[slot] = (double)A.dedq;
    else atomicAdd(&dedq_slot[slot], (double)A.dedq);   // zeroed by the launcher
    if (do_force & 1) {
      atomicAdd(F + 3 * (int64_t)slot, A.gx);
      atomicAdd(F + 3 * (int64_t)slot + 1, A.gy);
      atomicAdd(F + 3 * (int64_t)slot + 2, A.gz);
    }
    atomicAdd(&molacc[16 * m + 2], 0.5 * (double)A.ecc);
    atomicAdd(&molacc[16 * m + 3], 0.5 * (double)A.evdw);
    atomicAdd(&molacc[16 * m + 5], (double)A.dedq);
  }
}

// ---- table-driven pair kernel -------------------------------------------------------------------
// The pair energy is  e_ij = q_i q_j kappa(s) + w_p(s)  with s = r^2 (Angstrom^2) and p the element-pair type.  kappa and
// the w_p are tabulated on the host in float64 as cubic Hermite pieces on a grid that is uniform in the BITS of the fp32
// value of s: node k <-> s_k = float((k + kmin) << 17), i.e. 64 nodes per octave of s (1.1 % spacing in r at every
// distance: 640 nodes from r = 0.5 A to 16 A).  A piece holds (c0, c1, c2, c3) with f(s_k + u) = c0 + u (c1 + u (c2 + u c3));
// the force needs df/ds = c1 + u (2 c2 + 3 c3 u) only, because  d e_ij / d x_i = -2 (x_j - x_i) de/ds: no square root, no
// reciprocal, no exponential per pair (the analytic form costs ~110 instructions and 3 MUFU per pair, this one ~40).
// Interpolation error < 3e-7 relative (checked at build time against the analytic values at the piece mid-points).  The
// two pieces that contain a kink of the reference's potentials -- the ELU / DSF switch at Elu_Width (RawSymFunc.py:1353)
// and the end of the vdW polynomial switch (RawSymFunc.py:1391-1393), both continuous in value and slope only -- and
// anything closer than 0.5 A are evaluated analytically (pair_eval above).
#ifndef PT_MIN_CTAS
#define PT_MIN_CTAS 3
#endif
#define PT_SHIFT 17
#define PT_SMIN 0.25f
#define PT_SMAX 256.0f

struct PairTabMeta { int kmin, nnodes, nfn, kink_k, kink_v; };

static void pair_fn_host(const tm_params& p, int fn, const int* ei_ej, double s, double* f, double* dfds) {
  const double B = TM_BOHRPERA;
  const double r = sqrt(s);
  if (fn == 0) {      // kappa(R), R = B r  (TFCoulombEluSRDSFLR, RawSymFunc.py:1307-1359)
    const double R = B * r, alpha = p.dsf_alpha / B, Rl = p.ee_cutoff_off * B, Rs = p.elu_width * B;
    const double Zc = erfc(alpha * Rl) / Rl, Yc = 1.1283791671 * alpha * exp(-alpha * alpha * Rl * Rl) / Rl;
    double k = 0.0, dk = 0.0;
    if (R > Rs) {
      if (R <= Rl) {
        k = erfc(alpha * R) / R - Zc + (R - Rl) * (Zc / Rl + Yc);
        dk = -erfc(alpha * R) / (R * R) - M_2_SQRTPI * alpha * exp(-alpha * alpha * R * R) / R + (Zc / Rl + Yc);
      }
    } else {
      k = p.elu_alpha * (exp(R - Rs) - 1.0) + p.elu_shift;
      dk = p.elu_alpha * exp(R - Rs);
    }
    *f = k;
    *dfds = dk * B / (2.0 * r);
    return;
  }
  // w(R'), R' = B^2 r: the reference scales coordinates that are already in Bohr once more (TFVdwPolyLR, RawSymFunc.py:1377)
  const double Rp = B * B * r, pw = p.poly_width * B, t = Rp / pw;
  const double c6 = sqrt(p.C6[ei_ej[0]]) * sqrt(p.C6[ei_ej[1]]), Rsum = p.Rvdw[ei_ej[0]] + p.Rvdw[ei_ej[1]];
  double S = 1.0, dS = 0.0;
  if (t <= 0.0) { S = 0.0; }
  else if (t <= 1.0) { S = t * t * (3.0 - 2.0 * t); dS = 6.0 * t * (1.0 - t) / pw; }
  const double ff = c6 / pow(Rp, 6.0), X = 6.0 * pow(Rsum / Rp, 12.0), g = 1.0 / (1.0 + X);
  *f = -S * ff * g;
  const double dw = -(dS * ff * g + S * (-6.0 * ff / Rp) * g + S * ff * (12.0 * X * g * g / Rp));
  *dfds = dw * B * B / (2.0 * r);
}

static inline float pt_node_s(int k) {
  uint32_t b = (uint32_t)k << PT_SHIFT;
  float f;
  memcpy(&f, &b, 4);
  return f;
}
static inline int pt_node_of(float s) {
  uint32_t b;
  memcpy(&b, &s, 4);
  return (int)(b >> PT_SHIFT);
}

// (re)builds the device tables when the hyper-parameters changed; leaves pt_nfn = 0 when they would not fit
static int pair_tables(tm_ctx* c) {
  if (c->pairtab_gen == c->params_gen) return TM_OK;
  const tm_params& p = c->params;
  const int n_ele = c->desc.n_ele, n_elep = n_ele * (n_ele + 1) / 2, nfn = 1 + n_elep;
  const int kmin = pt_node_of(PT_SMIN), kmax = pt_node_of(PT_SMAX);
  const int nnodes = kmax - kmin;
  c->pt_nfn = 0;
  c->pairtab_gen = c->params_gen;
  if ((size_t)nnodes * nfn * 16 > 150 * 1024) return TM_OK;          // analytic kernel instead (more than 4 elements)
  if (p.ee_cutoff_off * p.ee_cutoff_off >= PT_SMAX * 0.999 || p.ee_cutoff_off < 1.0) return TM_OK;
  std::vector<float> tab((size_t)nnodes * nfn * 4);
  double worst = 0.0;
  for (int fn = 0; fn < nfn; fn++) {
    int ee[2] = {0, 0};
    if (fn > 0) {
      int l = 0;
      for (int i = 0; i < n_ele; i++)
        for (int j = i; j < n_ele; j++, l++)
          if (l == fn - 1) { ee[0] = i; ee[1] = j; }
    }
    const double rc2 = p.ee_cutoff_off * p.ee_cutoff_off, sk = p.elu_width * p.elu_width, rv0 = p.poly_width / TM_BOHRPERA, sv = rv0 * rv0;
    // error measure: |error| weighted by s = r^2 (the number of partners grows like r^2 dr) against the largest s |f|
    double scale = 1e-300;
    for (int k = 0; k <= nnodes; k++) {
      const double s0 = pt_node_s(k + kmin);
      double f0, d0;
      if (s0 > rc2) break;
      pair_fn_host(p, fn, ee, s0, &f0, &d0);
      scale = std::max(scale, s0 * fabs(f0));
    }
    for (int k = 0; k < nnodes; k++) {
      const double s0 = pt_node_s(k + kmin), s1 = pt_node_s(k + kmin + 1), h = s1 - s0;
      double f0, d0, f1, d1;
      pair_fn_host(p, fn, ee, s0, &f0, &d0);
      pair_fn_host(p, fn, ee, s1, &f1, &d1);
      const double c2 = (3.0 * (f1 - f0) / h - 2.0 * d0 - d1) / h, c3 = (2.0 * (f0 - f1) / h + d0 + d1) / (h * h);
      float* t = &tab[((size_t)fn * nnodes + k) * 4];
      t[0] = (float)f0; t[1] = (float)d0; t[2] = (float)c2; t[3] = (float)c3;
      // accuracy inside the piece (pieces with a kink are replaced by the analytic path, see the kernel)
      const bool kink = (fn == 0) ? (s0 < sk && sk <= s1) : (s0 < sv && sv <= s1);
      if (kink || s0 >= rc2) continue;
      for (int q = 1; q < 4; q++) {
        double fm, dm, u = 0.25 * q * h;
        pair_fn_host(p, fn, ee, s0 + u, &fm, &dm);
        worst = std::max(worst, (s0 + u) * fabs(f0 + u * (d0 + u * (c2 + u * c3)) - fm) / scale);
      }
    }
  }
  if (getenv("TM_TRACE")) fprintf(stderr, "[tm_trace] pair tables: %d nodes x %d functions, weighted interpolation error %.2e\n", nnodes, nfn, worst);
  if (!(worst < 1e-6)) return TM_OK;     // unusual parameters: keep the analytic kernel
  int rc;
  if ((rc = tm_buf(c, c->b_pairtab, tab.size() * 4))) return rc;
  TM_CUDA(cudaMemcpyAsync(c->b_pairtab.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, c->stream));
  TM_CUDA(cudaStreamSynchronize(c->stream));
  c->pt_kmin = kmin; c->pt_nnodes = nnodes; c->pt_nfn = nfn;
  c->pt_kink_k = pt_node_of((float)(p.elu_width * p.elu_width)) - kmin;
  const double rv = p.poly_width / TM_BOHRPERA;
  c->pt_kink_v = pt_node_of((float)(rv * rv)) - kmin;
  return TM_OK;
}

// Persistent CTAs (the tables are staged once per CTA); warps take centres round-robin.  Per centre the cell columns that
// can hold a partner are resolved one per lane (z-range from the centre's actual position: ONE contiguous run of the
// cell-sorted copy per column), the non-empty runs of 32 columns are laid end to end as one flat index space, and the
// warp walks that space 32 candidates at a time: fp32 distance test on the float4 record, then the table evaluation
// right there under the predicate (two thirds of the candidates are inside the sphere with the fine z bins of the
// lattice path, so compacting the survivors first would cost more than the idle lanes do).
template <bool ECC, bool VDW>
__global__ void __launch_bounds__(PAIR_WARPS * 32, PT_MIN_CTAS)
k_pair_tab(const SAtom* __restrict__ sat, const float4* __restrict__ pq, const uint8_t* __restrict__ pe, const int32_t* __restrict__ cstart,
           const GridParams* __restrict__ gp, const int32_t* __restrict__ rowsidx, const int32_t* __restrict__ rowslot, int64_t nrows,
           int64_t maxnatom, const __grid_constant__ DevParams P, const __grid_constant__ PairTabMeta M, const float4* __restrict__ tab_g,
           int do_force, float cutoff_A, int split, double* __restrict__ dedq_slot, float* __restrict__ F, double* __restrict__ molacc,
           const int32_t* __restrict__ rowmeta, int single_mol) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // (the wait is below, behind the table staging)
  extern __shared__ float4 s_tab[];                    // [nfn][nnodes]: a warp's 32 random nodes spread over all banks
  __shared__ int8_t s_fn[TM_MAX_ELE][TM_MAX_ELE];      // table of the vdW function of an element pair
  __shared__ double s_sum[PAIR_WARPS][3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < M.nnodes * M.nfn; i += blockDim.x) s_tab[i] = tab_g[i];
  if (threadIdx.x < TM_MAX_ELE * TM_MAX_ELE) s_fn[threadIdx.x / TM_MAX_ELE][threadIdx.x % TM_MAX_ELE] = (int8_t)(1 + P.pair_index[threadIdx.x / TM_MAX_ELE][threadIdx.x % TM_MAX_ELE]);
  asm volatile("griddepcontrol.wait;" ::: "memory");   // the tables are constants; everything else comes from the predecessors
  __syncthreads();
  const GridParams g = *gp;
  const float cell = (float)g.cell, icell = (float)g.inv_cell, izcell = (float)g.inv_zcell;
  const float rc2 = cutoff_A * cutoff_A;
  const float rcw = cutoff_A + P.skin, rcw2 = rcw * rcw;   // column walk: cutoff + Verlet skin (see k_pair)
  const unsigned lt_mask = (1u << lane) - 1u;
  const float w_img = (do_force & 2) ? 1.0f : 0.5f;    // see k_pair
  const int kmin = M.kmin, kink_k = M.kink_k, kink_v = M.kink_v, nnodes = M.nnodes;
  // rows in use (element ranges incl. their padding to the row tile): the allocation behind them is never visited
  const int64_t nwork = (int64_t)min((int64_t)rowmeta[2 * TM_MAX_ELE + 1], nrows) * split;
  double tE = 0.0, tV = 0.0, tD = 0.0;                  // single molecule: this warp's share of Ecc, Evdw, sum dE/dq
  for (int64_t gw = (int64_t)blockIdx.x * PAIR_WARPS + warp; gw < nwork; gw += (int64_t)gridDim.x * PAIR_WARPS) {
    const int64_t row = gw / split;
    const int sub = (int)(gw - row * split);
    const int slot = rowslot[row];
    if (slot < 0) continue;
    const int si = rowsidx[row];
    const SAtom ci = sat[si];
    const float4 pi = pq[si];
    const float qi = pi.w;
    const int ei = ci.e;
    const int m = (int)(slot / maxnatom);
    const float ox_ = (float)(ci.x - g.ox), oy_ = (float)(ci.y - g.oy), oz_ = (float)(ci.z - g.oz);
    int x0 = max(0, (int)floorf((ox_ - rcw) * icell)), x1 = min(g.gx - 1, (int)floorf((ox_ + rcw) * icell));
    int y0 = max(0, (int)floorf((oy_ - rcw) * icell)), y1 = min(g.gy - 1, (int)floorf((oy_ + rcw) * icell));
    const int ny = y1 - y0 + 1, ncol = (x1 - x0 + 1) * ny;
    const int8_t* fnrow = s_fn[ei];
    PairAcc A = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 * split < ncol; c0 += 32) {
      int cb = 0, len = 0;
      const int cidx = (c0 + lane) * split + sub;
      if (cidx < ncol) {
        int x = x0 + cidx / ny, y = y0 + cidx % ny;
        float lx = fmaxf(0.f, fmaxf(x * cell - ox_, ox_ - (x + 1) * cell));
        float ly = fmaxf(0.f, fmaxf(y * cell - oy_, oy_ - (y + 1) * cell));
        float rem = rcw2 - lx * lx - ly * ly;
        if (rem > 0.f) {
          float zr = sqrtf(rem);
          int z0 = max(0, (int)floorf((oz_ - zr) * izcell)), z1 = min(g.gz - 1, (int)floorf((oz_ + zr) * izcell));
          if (z1 >= z0) {
            int cbase = m * g.ncell_mol + (x * g.gy + y) * g.gz;
            cb = cstart[cbase + z0];
            len = cstart[cbase + z1 + 1] - cb;
          }
        }
      }
      // pack the non-empty runs into the low lanes, then lay them end to end: lane c holds run c as [endv - len, endv)
      const unsigned live = __ballot_sync(FULL, len > 0);
      const int nlive = __popc(live);
      if (nlive == 0) continue;
      {
        const int src = (lane < nlive) ? __fns(live, 0, lane + 1) : 0;
        const int cb2 = __shfl_sync(FULL, cb, src), len2 = __shfl_sync(FULL, len, src);
        cb = cb2;
        len = (lane < nlive) ? len2 : 0;
      }
      int endv = len;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(FULL, endv, o);
        if (lane >= o) endv += t;
      }
      const int T = __shfl_sync(FULL, endv, 31);
      const int offv = cb - (endv - len);          // j = flat index + offv
      // two independent 32-candidate groups per pass (lane: flat indices f0 + lane and f0 + 32 + lane): their loads and
      // table lookups overlap, which is what this latency-bound loop needs
      int ca = 0, cbb = 0;                          // runs of this lane's two flat indices (monotone over the passes)
      auto locate = [&](int idx, bool valid, int& c) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
          int e = __shfl_sync(FULL, endv, c);
          c += (valid && idx >= e) ? 1 : 0;
        }
        for (;;) {   // (the shuffle is evaluated by every lane: never put a *_sync intrinsic behind a short-circuit)
          const int e = __shfl_sync(FULL, endv, c);
          const bool more = valid && idx >= e;
          if (!__any_sync(FULL, more)) break;
          c += more ? 1 : 0;
        }
        return idx + __shfl_sync(FULL, offv, c);
      };
      auto evaluate = [&](const float4& d, float d2, int j) {
        const int pej = pe[j];
        const int ej = pej & 7;
        const float wj = (pej & 0x80) ? w_img : 1.0f;
        const uint32_t bits = __float_as_uint(d2);
        const int idn = (int)(bits >> PT_SHIFT) - kmin;
        if (idn < 0 || idn == kink_k || idn == kink_v) {       // closer than 0.5 A, or a piece with a kink: analytic
          pair_eval<ECC, VDW>(P, d2, d.x, d.y, d.z, qi, d.w, P.pk_c6[ei][ej], P.pk_rs12[ei][ej], wj, A);
        } else {
          const float u = d2 - __uint_as_float(bits & (0xffffffffu << PT_SHIFT));
          const float4* node = s_tab + idn;
          float deds = 0.f;
          if (ECC) {
            const float4 k = node[0];
            const float kap = fmaf(u, fmaf(u, fmaf(u, k.w, k.z), k.y), k.x);
            const float dk = fmaf(u, fmaf(u, 3.0f * k.w, 2.0f * k.z), k.y);
            const float qq = qi * d.w;
            A.ecc = fmaf(qq, kap, A.ecc);
            A.dedq = fmaf(d.w, kap, A.dedq);
            deds = qq * dk;
          }
          if (VDW) {
            const float4 v = node[fnrow[ej] * nnodes];
            A.evdw += fmaf(u, fmaf(u, fmaf(u, v.w, v.z), v.y), v.x);
            deds += fmaf(u, fmaf(u, 3.0f * v.w, 2.0f * v.z), v.y);
          }
          const float sc = -2.0f * wj * deds;          // d e / d x_i = de/ds * d s / d x_i = -2 (x_j - x_i) de/ds
          A.gx = fmaf(sc, d.x, A.gx); A.gy = fmaf(sc, d.y, A.gy); A.gz = fmaf(sc, d.z, A.gz);
        }
      };
      for (int f0 = 0; f0 < T; f0 += 64) {
        const int ia = f0 + lane, ib = ia + 32;
        const bool va = ia < T, vb = ib < T;
        if (cbb < ca) cbb = ca;
        const int ja = locate(ia, va, ca);
        const int jb = locate(ib, vb, cbb);
        float4 da = make_float4(0.f, 0.f, 0.f, 0.f), db = da;
        float d2a = 3.0e38f, d2b = 3.0e38f;
        if (va && ja != si) da = pq[ja];
        if (vb && jb != si) db = pq[jb];
        if (va && ja != si) {
          da.x -= pi.x; da.y -= pi.y; da.z -= pi.z;
          d2a = fmaf(da.x, da.x, fmaf(da.y, da.y, da.z * da.z));
        }
        if (vb && jb != si) {
          db.x -= pi.x; db.y -= pi.y; db.z -= pi.z;
          d2b = fmaf(db.x, db.x, fmaf(db.y, db.y, db.z * db.z));
        }
        if (d2a < rc2) evaluate(da, d2a, ja);
        if (d2b < rc2) evaluate(db, d2b, jb);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      A.ecc += __shfl_xor_sync(FULL, A.ecc, o);
      A.evdw += __shfl_xor_sync(FULL, A.evdw, o);
      A.dedq += __shfl_xor_sync(FULL, A.dedq, o);
      A.gx += __shfl_xor_sync(FULL, A.gx, o);
      A.gy += __shfl_xor_sync(FULL, A.gy, o);
      A.gz += __shfl_xor_sync(FULL, A.gz, o);
    }
    if (lane == 0) {
      if (split == 1) dedq_slot[slot] = (double)A.dedq;
      else atomicAdd(&dedq_slot[slot], (double)A.dedq);   // zeroed by the launcher
      if (do_force & 1) {
        atomicAdd(F + 3 * (int64_t)slot, A.gx);
        atomicAdd(F + 3 * (int64_t)slot + 1, A.gy);
        atomicAdd(F + 3 * (int64_t)slot + 2, A.gz);
      }
      if (single_mol) {
        tE += 0.5 * (double)A.ecc; tV += 0.5 * (double)A.evdw; tD += (double)A.dedq;
      } else {
        atomicAdd(&molacc[16 * m + 2], 0.5 * (double)A.ecc);
        atomicAdd(&molacc[16 * m + 3], 0.5 * (double)A.evdw);
        atomicAdd(&molacc[16 * m + 5], (double)A.dedq);
      }
    }
  }
  if (single_mol) {   // one set of atomics per CTA instead of one per centre (they all hit the same three addresses)
    if (lane == 0) { s_sum[warp][0] = tE; s_sum[warp][1] = tV; s_sum[warp][2] = tD; }
    __syncthreads();
    if (threadIdx.x < 3) {
      double t = 0.0;
      for (int w = 0; w < PAIR_WARPS; w++) t += s_sum[w][threadIdx.x];
      if (t != 0.0) atomicAdd(&molacc[threadIdx.x == 0 ? 2 : (threadIdx.x == 1 ? 3 : 5)], t);
    }
  }
}

int tm_launch_pair(tm_ctx* c, const SysView& s, int flags) {
  int rc;
  int64_t nq = s.periodic ? s.nreal : s.nslots;
  if ((rc = tm_buf(c, c->b_dedq, (size_t)nq * 8))) return rc;
  TM_CUDA(cudaMemsetAsync(c->b_dedq.p, 0, (size_t)nq * 8, c->stream));
  // one warp per centre; only really small systems (under ~1,000 centres) split a centre's columns over 2-8 warps.  (The
  // table kernel's persistent CTAs made the wider splitting of round 1 counterproductive — measured on B200, whole step:
  // 3,000 centres 0.2123 / 0.2135 / 0.2179 / 0.2316 ms with 1 / 2 / 4 / 8 warps per centre, a rank of 8 0.2257 / 0.2268 /
  // 0.2347 / 0.2414 ms, the 1,568-atom C5 box 0.5411 / 0.5440 / 0.5494 ms.)
  int64_t expect = std::max<int64_t>(1, s.slab_world > 1 ? (s.periodic ? s.nreal : s.nslots) / s.slab_world : s.ncent_max);
  int split = 1;
  while (split < 8 && expect * split < 1024) split *= 2;
  static const int split_env = getenv("TM_PAIR_SPLIT") ? atoi(getenv("TM_PAIR_SPLIT")) : 0;   // measurements
  if (split_env == 1 || split_env == 2 || split_env == 4 || split_env == 8) split = split_env;
  int blocks = (int)((s.nrows * split + PAIR_WARPS - 1) / PAIR_WARPS);
  if (nq > 0x7fffffff) { tm_set_error("too many slots"); return TM_EINVAL; }
  auto launch = [&](auto kern) {
    TM_LAUNCH(kern, blocks, PAIR_WARPS * 32, 0, c->stream, (const SAtom*)c->b_satom.p, (const float4*)c->b_qs.p, (const int32_t*)c->b_cstart.p,
                                                    (const GridParams*)c->b_grid.p, (const int32_t*)c->b_rowsidx.p, (const int32_t*)c->b_rowslot.p,
                                                    s.nrows, s.maxnatom, (int)nq, c->hp, ((flags & TM_F_FORCE) ? 1 : 0) | ((flags & TM_F_FOLD_IMAGES) ? 2 : 0),
                                                    (float)c->params.ee_cutoff_off, split, (double*)c->b_dedq.p, (float*)c->b_F.p, (double*)c->b_molacc.p);
  };
  bool ecc = c->hp.add_ecc != 0, vdw = (flags & TM_F_VDW) != 0;
  static const bool analytic_only = getenv("TM_PAIR_ANALYTIC") != nullptr;   // measurements / tests of the analytic kernel
  if ((rc = pair_tables(c))) return rc;
  if (c->pt_nfn > 0 && !analytic_only && (ecc || vdw)) {
    PairTabMeta M{c->pt_kmin, c->pt_nnodes, c->pt_nfn, c->pt_kink_k, c->pt_kink_v};
    const size_t smem = (size_t)M.nnodes * M.nfn * 16;
    auto launch_tab = [&](auto kern) -> int {
      // per kernel AND device (the three instantiations share this lambda body: same pointer type)
      static std::map<std::pair<const void*, int>, size_t> conf;
      size_t& have = conf[std::make_pair((const void*)kern, c->device)];
      if (smem > have) {
        TM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        have = smem;
      }
      int occ = 1, sms = 148;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PAIR_WARPS * 32, smem);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
      int grid = std::max(1, std::min(blocks, sms * std::max(1, occ)));
      TM_LAUNCH(kern, grid, PAIR_WARPS * 32, smem, c->stream, (const SAtom*)c->b_satom.p, (const float4*)c->b_qs.p, (const uint8_t*)c->b_pe.p,
                                                      (const int32_t*)c->b_cstart.p, (const GridParams*)c->b_grid.p, (const int32_t*)c->b_rowsidx.p,
                                                      (const int32_t*)c->b_rowslot.p, s.nrows, s.maxnatom, c->hp, M, (const float4*)c->b_pairtab.p,
                                                      ((flags & TM_F_FORCE) ? 1 : 0) | ((flags & TM_F_FOLD_IMAGES) ? 2 : 0), (float)c->params.ee_cutoff_off,
                                                      split, (double*)c->b_dedq.p, (float*)c->b_F.p, (double*)c->b_molacc.p,
                                                      (const int32_t*)c->b_rowmeta.p, s.nmol == 1 ? 1 : 0);
      c->launches++;
      TM_CUDA(cudaGetLastError());
      return TM_OK;
    };
    if (ecc && vdw) return launch_tab(k_pair_tab<true, true>);
    if (ecc) return launch_tab(k_pair_tab<true, false>);
    return launch_tab(k_pair_tab<false, true>);
  }
  if (ecc && vdw) launch(k_pair<true, true>);
  else if (ecc) launch(k_pair<true, false>);
  else if (vdw) launch(k_pair<false, true>);
  else launch(k_pair<false, false>);
  c->launches++;
  TM_CUDA(cudaGetLastError());
  return TM_OK;
}
