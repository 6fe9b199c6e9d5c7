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

#define FULL 0xffffffffu
#define PAIR_WARPS 8
#define QCAP 96

// ---- charges -----------------------------------------------------------------------------------
// y_charge[row] -> qraw_slot[slot]; per-molecule sums in double
__global__ void k_qraw_scatter(const float* __restrict__ y, const int32_t* __restrict__ rowslot, int64_t nrows, double* __restrict__ qraw_slot) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x) {
    int s = rowslot[r];
    if (s >= 0) qraw_slot[s] = (double)y[r];
  }
}

// molsum[m] = sum over real slots of molecule m (one block per molecule chunk)
__global__ void k_mol_sum(const double* __restrict__ v, const int32_t* __restrict__ Z, int64_t maxnatom, int64_t nvalid_per_mol, double* __restrict__ molacc, int field) {
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
                          const GridParams* __restrict__ gp, int64_t nreal, int periodic, float4* __restrict__ pq) {
  int m = blockIdx.y;
  double mean = molacc[16 * m + 4] * inv_n[m];
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
  for (int i = ib + blockIdx.x * blockDim.x + threadIdx.x; i < ie; i += gridDim.x * blockDim.x) {
    SAtom a = sat[i];
    int64_t s = a.slot;
    if (periodic) s = s % nreal;
    pq[i] = make_float4((float)(a.x - g.ox), (float)(a.y - g.oy), (float)(a.z - g.oz), (float)(qraw_slot[s] - mean));
  }
}

int tm_launch_charges(tm_ctx* c, const SysView& s) {
  int rc;
  int64_t nq = s.periodic ? s.nreal : s.nslots;           // slots that carry an own charge
  int64_t nq_per_mol = s.periodic ? s.nreal : s.maxnatom;
  if ((rc = tm_buf(c, c->b_q, (size_t)nq * 8 * 2))) return rc;  // [qraw_slot | q_slot]
  if ((rc = tm_buf(c, c->b_qs, (size_t)s.nslots * 16))) return rc;
  double* qraw = (double*)c->b_q.p;
  double* q = qraw + nq;
  // molacc (zeroed by the caller at the start of the evaluation), stride 16 doubles per molecule:
  //   [0] Etotal [1] Ebp [2] Ecc [3] Evdw [4] sum q_raw [5] sum dE/dq [6..8] dipole
  double* molacc = (double*)c->b_molacc.p;
  // slab mode: qraw_slot was already combined over the ranks and written into b_q; otherwise the tensor-core forward
  // pass has scattered q_raw and summed it per molecule (k_y_reduce), and the fp32 mode does both here
  const bool fused = c->y_fused && !s.slab_api;
  if (!fused) {
    if (!s.slab_api) {
      int blocks = (int)((s.nrows + 255) / 256);
      TM_CUDA(cudaMemsetAsync(qraw, 0, (size_t)nq * 8, c->stream));
      k_qraw_scatter<<<blocks, 256, 0, c->stream>>>((const float*)c->b_y[TM_NET_CHARGE].p, (const int32_t*)c->b_rowslot.p, s.nrows, qraw);
      c->launches++;
    }
    dim3 gms((unsigned)s.nmol, (unsigned)std::max<int64_t>(1, std::min<int64_t>((nq_per_mol + 2047) / 2048, 64)));
    k_mol_sum<<<gms, 256, 0, c->stream>>>(qraw, (const int32_t*)c->b_Z.p, s.maxnatom, nq_per_mol, molacc, 4);
    c->launches++;
  }
  int64_t per_mol = std::max<int64_t>(nq_per_mol, s.nslots / std::max<int64_t>(1, s.nmol));
  dim3 g((unsigned)std::max<int64_t>(1, std::min<int64_t>((per_mol + 255) / 256, 148 * 8)), (unsigned)s.nmol);
  k_charges<<<g, 256, 0, c->stream>>>(qraw, molacc, (const double*)c->b_natom.p, (const double*)c->b_pos.p, (const int32_t*)c->b_Z.p, s.maxnatom,
                                      nq_per_mol, q, (const SAtom*)c->b_satom.p, (const int32_t*)c->b_cstart.p, (const GridParams*)c->b_grid.p, s.nreal,
                                      s.periodic, (float4*)c->b_qs.p);
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
  float cell = (float)g.cell, icell = (float)g.inv_cell;
  float rc2 = cutoff_A * cutoff_A;
  const unsigned lt_mask = (1u << lane) - 1u;
  // cell columns that can hold a partner, from the centre's actual position (not its cell): |dx| <= rc
  int x0 = max(0, (int)floorf((pi.x - cutoff_A) * icell)), x1 = min(g.gx - 1, (int)floorf((pi.x + cutoff_A) * icell));
  int y0 = max(0, (int)floorf((pi.y - cutoff_A) * icell)), y1 = min(g.gy - 1, (int)floorf((pi.y + cutoff_A) * icell));
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
      float lx = fmaxf(0.f, fmaxf(x * cell - pi.x, pi.x - (x + 1) * cell));   // distance from the centre to the column's slab in x
      float ly = fmaxf(0.f, fmaxf(y * cell - pi.y, pi.y - (y + 1) * cell));
      float rem = rc2 - lx * lx - ly * ly;
      if (rem > 0.f) {
        float zr = sqrtf(rem);
        int z0 = max(0, (int)floorf((pi.z - zr) * icell)), z1 = min(g.gz - 1, (int)floorf((pi.z + zr) * icell));
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

int tm_launch_pair(tm_ctx* c, const SysView& s, int flags) {
  int rc;
  int64_t nq = s.periodic ? s.nreal : s.nslots;
  if ((rc = tm_buf(c, c->b_dedq, (size_t)nq * 8))) return rc;
  TM_CUDA(cudaMemsetAsync(c->b_dedq.p, 0, (size_t)nq * 8, c->stream));
  // enough warps to fill the machine: one per centre for large systems, up to 8 per centre for small ones / slabs
  int64_t expect = std::max<int64_t>(1, s.slab_world > 1 ? (s.periodic ? s.nreal : s.nslots) / s.slab_world : s.ncent_max);
  int split = 1;
  while (split < 8 && expect * split < 148 * 48) split *= 2;
  int blocks = (int)((s.nrows * split + PAIR_WARPS - 1) / PAIR_WARPS);
  if (nq > 0x7fffffff) { tm_set_error("too many slots"); return TM_EINVAL; }
  auto launch = [&](auto kern) {
    kern<<<blocks, PAIR_WARPS * 32, 0, c->stream>>>((const SAtom*)c->b_satom.p, (const float4*)c->b_qs.p, (const int32_t*)c->b_cstart.p,
                                                    (const GridParams*)c->b_grid.p, (const int32_t*)c->b_rowsidx.p, (const int32_t*)c->b_rowslot.p,
                                                    s.nrows, s.maxnatom, (int)nq, c->hp, ((flags & TM_F_FORCE) ? 1 : 0) | ((flags & TM_F_FOLD_IMAGES) ? 2 : 0),
                                                    (float)c->params.ee_cutoff_off, split, (double*)c->b_dedq.p, (float*)c->b_F.p, (double*)c->b_molacc.p);
  };
  bool ecc = c->hp.add_ecc != 0, vdw = (flags & TM_F_VDW) != 0;
  if (ecc && vdw) launch(k_pair<true, true>);
  else if (ecc) launch(k_pair<true, false>);
  else if (vdw) launch(k_pair<false, true>);
  else launch(k_pair<false, false>);
  c->launches++;
  TM_CUDA(cudaGetLastError());
  return TM_OK;
}
