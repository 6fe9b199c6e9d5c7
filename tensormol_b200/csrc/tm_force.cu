// K6: fused descriptor-gradient + force scatter-add.  Replaces the part of
// tf.gradients(Etotal, xyzs) (TFMolInstanceDirect.py:5761 / 5999) that flows through the symmetry
// functions: given A[row, :] = dE/dG_row (energy net) + u_row * dq_raw/dG_row (charge net, u_row =
// dE/dq_raw,row), accumulate  dE/dx_i, dE/dx_j, dE/dx_k  for every pair and triple of every centre.
//
// Reference gradient convention: every row of the (real+image) coordinate array is an
// independent variable and only rows < nreal are returned (TFMolManage.py:1353), so contributions
// addressed to image rows are DROPPED (unless TM_F_FOLD_IMAGES asks to fold them onto slot % nreal).
//
// Mapping: one warp per centre row; A row staged in shared memory with (n+1)-padded channel strides.
//   radial : lanes = neighbours, each lane loops the nRs_r Gaussians of its pair
//   angular: lanes = triples, each lane contracts the 64-wide channel block of its triple in registers
// Neighbour forces of the angular part are combined in a per-warp shared tile (shared atomics) and
// flushed with one global atomic per neighbour component.
#include "tm_internal.h"
#include <cstdlib>

#define FULL 0xffffffffu
#define FORCE_WARPS 8

__device__ __forceinline__ void tri_inv_f(int t, int& j, int& k) {
  k = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)t)) * 0.5f);
  while (k * (k - 1) / 2 > t) k--;
  while ((k + 1) * k / 2 <= t) k++;
  j = t - k * (k - 1) / 2;
}

size_t tm_force_smem_floats_per_warp(const DevParams& P) {
  return (size_t)P.n_ele * (P.nRs_r + 1) + (size_t)P.n_elep * (P.nsym + 1) + 6 * TM_ANG_CAP + 2 * TM_ANG_CAP + 3 * TM_ANG_CAP;
}

__device__ __forceinline__ float f_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float f_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

template <int NAS_MAX, int NRS_MAX>
__global__ void __launch_bounds__(FORCE_WARPS * 32)
k_force(const SAtom* __restrict__ sat, const int32_t* __restrict__ rowsidx, const int32_t* __restrict__ rowslot,
        const int32_t* __restrict__ nbcnt, const uint32_t* __restrict__ nbr, int64_t nrows, const __grid_constant__ DevParams P,
        const float* __restrict__ dGe, const float* __restrict__ dGq, const double* __restrict__ dedq_slot, const double* __restrict__ molacc,
        const double* __restrict__ inv_n, int64_t maxnatom, int64_t nreal_slots, int fold, float* __restrict__ F, int wfloats) {
  TM_PDL_PROLOGUE;
  extern __shared__ float smem[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t row = (int64_t)blockIdx.x * FORCE_WARPS + warp;
  if (row >= nrows) return;
  int slot = rowslot[row];
  if (slot < 0) return;
  float* ws = smem + (size_t)warp * wfloats;
  const int rstr = P.nRs_r + 1, astr = P.nsym + 1;
  float* Ar = ws;                                // [n_ele][nRs_r+1]
  float* Aa = Ar + P.n_ele * rstr;               // [n_elep][nsym+1]
  float* ax = Aa + P.n_elep * astr;
  float* ay = ax + TM_ANG_CAP;
  float* az = ay + TM_ANG_CAP;
  float* ar = az + TM_ANG_CAP;
  float* afc = ar + TM_ANG_CAP;
  float* adfc = afc + TM_ANG_CAP;
  int* ae = (int*)(adfc + TM_ANG_CAP);
  int* aslot = ae + TM_ANG_CAP;                  // destination slot for the force, -1 = dropped
  float* Ft = (float*)(aslot + TM_ANG_CAP);      // [ANG_CAP][3]

  // stage A = dGe + u*dGq
  {
    const float* ge = dGe + row * P.Dp;
    const float* gq = dGq + row * P.Dp;
    // u = dE/dq_raw = dE/dq_slot - mean_mol(dE/dq): backward of the neutralisation (TFMolInstanceDirect.py:5274-5277)
    float uu = 0.f;
    if (P.add_ecc) {
      int m = (int)(slot / maxnatom);
      uu = (float)(dedq_slot[slot] - molacc[16 * m + 5] * inv_n[m]);
    }
    int nrad = P.n_ele * P.nRs_r;
    for (int i = lane; i < nrad; i += 32) {
      int q = i / P.nRs_r, s = i - q * P.nRs_r;
      Ar[q * rstr + s] = ge[i] + uu * gq[i];
    }
    int na = P.n_elep * P.nsym;
    for (int i = lane; i < na; i += 32) {
      int p = i / P.nsym, s = i - p * P.nsym;
      Aa[p * astr + s] = ge[nrad + i] + uu * gq[nrad + i];
    }
    for (int i = lane; i < 3 * TM_ANG_CAP; i += 32) Ft[i] = 0.f;
  }
  __syncwarp();

  SAtom ci = sat[rowsidx[row]];
  int b = (int)row * TM_NB_STRIDE, e = b + nbcnt[row];
  const float nel2 = -P.eta * 1.4426950408889634f;   // exp(-eta x) = 2^(nel2 x)
  float gix = 0.f, giy = 0.f, giz = 0.f;   // dE/dx_i accumulated by this lane
  int nang = 0;
  for (int j0 = b; j0 < e; j0 += 32) {
    int j = j0 + lane;
    bool isang = false;
    float dx = 0.f, dy = 0.f, dz = 0.f, r = 1.f;
    int ej = 0, dst = -1;
    if (j < e) {
      uint32_t en = nbr[j];
      isang = (en >> 31) != 0;
      SAtom a = sat[en & 0x7fffffffu];
      dx = (float)(a.x - ci.x);
      dy = (float)(a.y - ci.y);
      dz = (float)(a.z - ci.z);
      r = sqrtf(dx * dx + dy * dy + dz * dz);
      ej = a.e;
      dst = (a.slot < nreal_slots) ? a.slot : (fold ? (int)(a.slot % nreal_slots) : -1);
      // radial: dE/dr = sum_s A[e_j][s] * d/dr [ exp(-eta (r-Rs)^2) fc(r) ]
      float arg = P.pi_over_rRc * r;
      float sn, cs;
      __sincosf(arg, &sn, &cs);                     // arg in [0, pi]: absolute error < 5e-7
      float fc = 0.5f * (cs + 1.0f);
      float dfc = -0.5f * sn * P.pi_over_rRc;
      if (P.skin_on) {   // the rows reach out to cutoff + skin (tm_set_skin): the cutoffs are applied here
        if (!(r < P.r_Rc)) { fc = 0.f; dfc = 0.f; }
        isang = isang && (r < P.a_Rc);
      }
      const float* Arow = Ar + ej * rstr;
      float dEdr = 0.f;
      const float m2ef = -2.0f * P.eta * fc;
      for (int s = 0; s < P.nRs_r; s++) {
        float d = r - P.Rs_r[s];
        float g = f_ex2(nel2 * d * d);               // exp(-eta d^2)
        dEdr = fmaf(Arow[s] * g, fmaf(m2ef, d, dfc), dEdr);
      }
      float sc = dEdr * f_rcp(r);
      float gx = sc * dx, gy = sc * dy, gz = sc * dz;   // dE/dx_j  (d = x_j - x_i)
      gix -= gx; giy -= gy; giz -= gz;
      if (dst >= 0) {
        atomicAdd(F + 3 * (int64_t)dst, gx);
        atomicAdd(F + 3 * (int64_t)dst + 1, gy);
        atomicAdd(F + 3 * (int64_t)dst + 2, gz);
      }
    }
    unsigned mk = __ballot_sync(FULL, isang);
    if (isang) {
      int pos = nang + __popc(mk & ((1u << lane) - 1));
      if (pos < TM_ANG_CAP) {
        ax[pos] = dx; ay[pos] = dy; az[pos] = dz; ar[pos] = r;
        float sa_, ca_;
        __sincosf(P.pi_over_aRc * r, &sa_, &ca_);
        afc[pos] = 0.5f * (ca_ + 1.0f);              // fc(r, Ra) and its derivative, once per neighbour
        adfc[pos] = -0.5f * sa_ * P.pi_over_aRc;
        ae[pos] = ej;
        aslot[pos] = dst;
      }
    }
    nang += __popc(mk);
  }
  nang = min(nang, TM_ANG_CAP);
  __syncwarp();

  int ntrip = nang * (nang - 1) / 2;
  for (int t0 = 0; t0 < ntrip; t0 += 32) {
    int t = t0 + lane;
    if (t < ntrip) {
      int j, k;
      tri_inv_f(t, j, k);
      float ajx = ax[j], ajy = ay[j], ajz = az[j], akx = ax[k], aky = ay[k], akz = az[k];
      float ra = ar[j], rb = ar[k];
      float ira = f_rcp(ra), irb = f_rcp(rb);
      // unit vectors
      float uax = ajx * ira, uay = ajy * ira, uaz = ajz * ira;
      float ubx = akx * irb, uby = aky * irb, ubz = akz * irb;
      float c = uax * ubx + uay * uby + uaz * ubz;
      float nx = uay * ubz - uaz * uby, ny = uaz * ubx - uax * ubz, nz = uax * uby - uay * ubx;
      float s = sqrtf(nx * nx + ny * ny + nz * nz);
      c = fminf(1.0f, fmaxf(-1.0f, c));
      float fa = afc[j], fb = afc[k];
      float dfa = adfc[j], dfb = adfc[k];
      float rho = 0.5f * (ra + rb);
      float E[NRS_MAX], dE[NRS_MAX];
#pragma unroll
      for (int q = 0; q < NRS_MAX; q++) {
        if (q < P.nRs_a) {
          float d = rho - P.Rs_a[q];
          float g = f_ex2(nel2 * d * d);
          E[q] = g;
          dE[q] = -2.0f * P.eta * d * g;
        } else {
          E[q] = 0.f; dE[q] = 0.f;
        }
      }
      int p = P.pair_index[ae[j]][ae[k]];
      const float* Ap = Aa + p * astr;
      float W = 0.f, Wt = 0.f, Wr = 0.f;
#pragma unroll
      for (int a = 0; a < NAS_MAX; a++) {
        if (a < P.nAs) {
          float ca = P.cosA[a], sa = P.sinA[a];
          float base = fmaxf(1.0f + c * ca + s * sa, 0.f);
          float T, dT;   // T = pref*base^zeta ; dT = dT/dtheta = -zeta*pref*base^(zeta-1) * sin(theta-theta_a)
          float sind = s * ca - c * sa;
          if (P.zeta_is8) {
            float b2 = base * base, b4 = b2 * b2;
            float b7 = b4 * b2 * base;
            T = P.zeta_pref * b7 * base;
            dT = -8.0f * P.zeta_pref * b7 * sind;
          } else {
            float bm = powf(base, P.zeta - 1.0f);
            T = P.zeta_pref * bm * base;
            dT = -P.zeta * P.zeta_pref * bm * sind;
          }
          float ua = 0.f, va = 0.f;
#pragma unroll
          for (int q = 0; q < NRS_MAX; q++) {
            if (q < P.nRs_a) {
              float Aas = Ap[a * P.nRs_a + q];
              ua += Aas * E[q];
              va += Aas * dE[q];
            }
          }
          W += T * ua;
          Wt += dT * ua;
          Wr += T * va;
        }
      }
      // V = W fa fb ; gradients w.r.t. a = x_j - x_i and b = x_k - x_i
      float ff = fa * fb;
      float dVdt = Wt * ff;
      float dVda_r = 0.5f * Wr * ff + W * dfa * fb;   // along a_hat
      float dVdb_r = 0.5f * Wr * ff + W * fa * dfb;   // along b_hat
      float is = (s > 1e-6f) ? f_rcp(s) : 0.f;
      // dtheta/da = -(b_hat - c a_hat)/(|a| s) ; dtheta/db = -(a_hat - c b_hat)/(|b| s)
      float ta = -dVdt * is * ira, tb = -dVdt * is * irb;
      float gax = ta * (ubx - c * uax) + dVda_r * uax;
      float gay = ta * (uby - c * uay) + dVda_r * uay;
      float gaz = ta * (ubz - c * uaz) + dVda_r * uaz;
      float gbx = tb * (uax - c * ubx) + dVdb_r * ubx;
      float gby = tb * (uay - c * uby) + dVdb_r * uby;
      float gbz = tb * (uaz - c * ubz) + dVdb_r * ubz;
      gix -= gax + gbx; giy -= gay + gby; giz -= gaz + gbz;
      atomicAdd(&Ft[3 * j], gax); atomicAdd(&Ft[3 * j + 1], gay); atomicAdd(&Ft[3 * j + 2], gaz);
      atomicAdd(&Ft[3 * k], gbx); atomicAdd(&Ft[3 * k + 1], gby); atomicAdd(&Ft[3 * k + 2], gbz);
    }
  }
  __syncwarp();
  for (int i = lane; i < nang; i += 32) {
    int dst = aslot[i];
    if (dst >= 0) {
      atomicAdd(F + 3 * (int64_t)dst, Ft[3 * i]);
      atomicAdd(F + 3 * (int64_t)dst + 1, Ft[3 * i + 1]);
      atomicAdd(F + 3 * (int64_t)dst + 2, Ft[3 * i + 2]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    gix += __shfl_xor_sync(FULL, gix, o);
    giy += __shfl_xor_sync(FULL, giy, o);
    giz += __shfl_xor_sync(FULL, giz, o);
  }
  if (lane == 0) {
    atomicAdd(F + 3 * (int64_t)slot, gix);
    atomicAdd(F + 3 * (int64_t)slot + 1, giy);
    atomicAdd(F + 3 * (int64_t)slot + 2, giz);
  }
}

// ---- fast path: the ANI-1 default grid (8 x 8 angular functions per pair channel, 32 radial functions) ----
// Same arithmetic as k_force; what changes is how the work is issued (k_force: 3.3k warp instructions per water centre,
// a fifth of them and 45 % of the stall samples in the compare-and-swap loops of the shared-memory float atomics):
//   * the per-neighbour forces of the angular part are no longer combined with shared atomics: every triple writes its
//     two gradient vectors to a per-round tile, and lane n then sums the entries of neighbour n (the triples (j, k) with
//     k = n are a contiguous run of the triangular enumeration, those with j = n sit one per k-row);
//   * A = dGe + u dGq is staged with 128-bit loads and shift index arithmetic, the angular rows with a 16-byte aligned
//     pitch (68) so that the contraction reads them with 128-bit loads;
//   * a last radial chunk with few neighbours (water: 8 of 40) is spread over the idle lanes, 2..32 lanes per neighbour
//     taking interleaved Gaussians.
#define FF_RSTR 33
#define FF_ASTR 68
size_t tm_force_fast_smem_floats_per_warp(const DevParams& P) {
  size_t n = (size_t)P.n_elep * FF_ASTR + (size_t)P.n_ele * FF_RSTR + 32 + 6 * TM_ANG_CAP + 2 * TM_ANG_CAP + 2 * 32 * 4;
  return (n + 3) / 4 * 4;
}

__global__ void __launch_bounds__(FORCE_WARPS * 32, 3)
k_force_fast(const SAtom* __restrict__ sat, const int32_t* __restrict__ rowsidx, const int32_t* __restrict__ rowslot,
             const int32_t* __restrict__ nbcnt, const uint32_t* __restrict__ nbr, int64_t nrows, const __grid_constant__ DevParams P,
             const float* __restrict__ dGe, const float* __restrict__ dGq, const double* __restrict__ dedq_slot, const double* __restrict__ molacc,
             const double* __restrict__ inv_n, int64_t maxnatom, int64_t nreal_slots, int fold, float* __restrict__ F, int wfloats) {
  TM_PDL_PROLOGUE;
  constexpr int NA = 8, NR = 8, NSYM = 64, NRAD = 32;
  extern __shared__ float smem[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t row = (int64_t)blockIdx.x * FORCE_WARPS + warp;
  if (row >= nrows) return;
  int slot = rowslot[row];
  if (slot < 0) return;
  float* ws = smem + (size_t)warp * wfloats;
  float* Aa = ws;                                // [n_elep][FF_ASTR], rows 16-byte aligned
  float4* GA = (float4*)(Aa + P.n_elep * FF_ASTR);   // [32] dV/da of this round's triples
  float4* GB = GA + 32;                          // [32] dV/db
  float* Ar = (float*)(GB + 32);                 // [n_ele][FF_RSTR]
  float* Rs = Ar + P.n_ele * FF_RSTR;            // [32] copy of Rs_r for lane-dependent indexing
  float* ax = Rs + 32;
  float* ay = ax + TM_ANG_CAP;
  float* az = ay + TM_ANG_CAP;
  float* ar = az + TM_ANG_CAP;
  float* afc = ar + TM_ANG_CAP;
  float* adfc = afc + TM_ANG_CAP;
  int* ae = (int*)(adfc + TM_ANG_CAP);
  int* aslot = ae + TM_ANG_CAP;                  // destination slot for the force, -1 = dropped

  // stage A = dGe + u*dGq
  {
    // u = dE/dq_raw = dE/dq_slot - mean_mol(dE/dq): backward of the neutralisation (TFMolInstanceDirect.py:5274-5277)
    float uu = 0.f;
    if (P.add_ecc) {
      int m = (int)(slot / maxnatom);
      uu = (float)(dedq_slot[slot] - molacc[16 * m + 5] * inv_n[m]);
    }
    const float4* ge = reinterpret_cast<const float4*>(dGe + row * P.Dp);
    const float4* gq = reinterpret_cast<const float4*>(dGq + row * P.Dp);
    const int nrad4 = P.n_ele * (NRAD / 4), n4 = nrad4 + P.n_elep * (NSYM / 4);
    for (int i4 = lane; i4 < n4; i4 += 32) {
      float4 a = ge[i4], q = gq[i4];
      float4 v = make_float4(fmaf(uu, q.x, a.x), fmaf(uu, q.y, a.y), fmaf(uu, q.z, a.z), fmaf(uu, q.w, a.w));
      if (i4 < nrad4) {
        float* d = Ar + (i4 >> 3) * FF_RSTR + 4 * (i4 & 7);
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      } else {
        int j4 = i4 - nrad4;
        *reinterpret_cast<float4*>(Aa + (j4 >> 4) * FF_ASTR + 4 * (j4 & 15)) = v;
      }
    }
    Rs[lane] = P.Rs_r[lane];
  }
  __syncwarp();

  SAtom ci = sat[rowsidx[row]];
  int b = (int)row * TM_NB_STRIDE, e = b + nbcnt[row];
  const float nel2 = -P.eta * 1.4426950408889634f;   // exp(-eta x) = 2^(nel2 x)
  float gix = 0.f, giy = 0.f, giz = 0.f;   // dE/dx_i accumulated by this lane
  int nang = 0;
  for (int j0 = b; j0 < e;) {
    // lanes per neighbour in this chunk: 1 for a full chunk, else as many as fit (interleaved Gaussians)
    const int remaining = e - j0;
    int lsub = 0;
    while (lsub < 5 && (remaining << (lsub + 1)) <= 32) lsub++;
    const int sub = 1 << lsub;
    const int cnt = min(remaining, 32);
    const int nb = lane >> lsub, part = lane & (sub - 1);
    const bool valid = nb < cnt;
    bool isang = false;
    float dx = 0.f, dy = 0.f, dz = 0.f, r = 1.f;
    int ej = 0, dst = -1;
    float fc = 0.f, dfc = 0.f;
    if (valid) {
      uint32_t en = nbr[j0 + nb];
      isang = (en >> 31) != 0 && part == 0;
      SAtom a = sat[en & 0x7fffffffu];
      dx = (float)(a.x - ci.x);
      dy = (float)(a.y - ci.y);
      dz = (float)(a.z - ci.z);
      r = sqrtf(dx * dx + dy * dy + dz * dz);
      ej = a.e;
      dst = (a.slot < nreal_slots) ? a.slot : (fold ? (int)(a.slot % nreal_slots) : -1);
      float sn, cs;
      __sincosf(P.pi_over_rRc * r, &sn, &cs);      // arg in [0, pi]: absolute error < 5e-7
      fc = 0.5f * (cs + 1.0f);
      dfc = -0.5f * sn * P.pi_over_rRc;
      if (P.skin_on) {   // the rows reach out to cutoff + skin (tm_set_skin): the cutoffs are applied here
        if (!(r < P.r_Rc)) { fc = 0.f; dfc = 0.f; }
        isang = isang && (r < P.a_Rc);
      }
    }
    // radial: dE/dr = sum_s A[e_j][s] * d/dr [ exp(-eta (r-Rs)^2) fc(r) ]
    const float* Arow = Ar + ej * FF_RSTR;
    const float m2ef = -2.0f * P.eta * fc;
    float dEdr = 0.f;
    if (sub == 1) {
#pragma unroll 8
      for (int s = 0; s < NRAD; s++) {
        float d = r - P.Rs_r[s];
        float g = f_ex2(nel2 * d * d);               // exp(-eta d^2)
        dEdr = fmaf(Arow[s] * g, fmaf(m2ef, d, dfc), dEdr);
      }
    } else {
      for (int s = part; s < NRAD; s += sub) {
        float d = r - Rs[s];
        float g = f_ex2(nel2 * d * d);
        dEdr = fmaf(Arow[s] * g, fmaf(m2ef, d, dfc), dEdr);
      }
      for (int o = 1; o < sub; o <<= 1) dEdr += __shfl_xor_sync(FULL, dEdr, o);
    }
    if (valid && part == 0) {
      float sc = dEdr * f_rcp(r);
      float gx = sc * dx, gy = sc * dy, gz = sc * dz;   // dE/dx_j  (d = x_j - x_i)
      gix -= gx; giy -= gy; giz -= gz;
      if (dst >= 0) {
        atomicAdd(F + 3 * (int64_t)dst, gx);
        atomicAdd(F + 3 * (int64_t)dst + 1, gy);
        atomicAdd(F + 3 * (int64_t)dst + 2, gz);
      }
    }
    unsigned mk = __ballot_sync(FULL, isang);
    if (isang) {
      int pos = nang + __popc(mk & ((1u << lane) - 1));
      if (pos < TM_ANG_CAP) {
        ax[pos] = dx; ay[pos] = dy; az[pos] = dz; ar[pos] = r;
        float sa_, ca_;
        __sincosf(P.pi_over_aRc * r, &sa_, &ca_);
        afc[pos] = 0.5f * (ca_ + 1.0f);              // fc(r, Ra) and its derivative, once per neighbour
        adfc[pos] = -0.5f * sa_ * P.pi_over_aRc;
        ae[pos] = ej;
        aslot[pos] = dst;
      }
    }
    nang += __popc(mk);
    j0 += cnt;
  }
  nang = min(nang, TM_ANG_CAP);
  __syncwarp();

  // forces on the angular neighbours n = lane and n = lane + 32, accumulated over the rounds
  float f0x = 0.f, f0y = 0.f, f0z = 0.f, f1x = 0.f, f1y = 0.f, f1z = 0.f;
  int ntrip = nang * (nang - 1) / 2;
  for (int t0 = 0; t0 < ntrip; t0 += 32) {
    int t = t0 + lane;
    const int cnt = min(32, ntrip - t0);
    int j = 0, k = 1;
    tri_inv_f(min(t, ntrip - 1), j, k);
    if (t < ntrip) {
      float ajx = ax[j], ajy = ay[j], ajz = az[j], akx = ax[k], aky = ay[k], akz = az[k];
      float ra = ar[j], rb = ar[k];
      float ira = f_rcp(ra), irb = f_rcp(rb);
      // unit vectors
      float uax = ajx * ira, uay = ajy * ira, uaz = ajz * ira;
      float ubx = akx * irb, uby = aky * irb, ubz = akz * irb;
      float c = uax * ubx + uay * uby + uaz * ubz;
      float nx = uay * ubz - uaz * uby, ny = uaz * ubx - uax * ubz, nz = uax * uby - uay * ubx;
      float s = sqrtf(nx * nx + ny * ny + nz * nz);
      c = fminf(1.0f, fmaxf(-1.0f, c));
      float fa = afc[j], fb = afc[k];
      float dfa = adfc[j], dfb = adfc[k];
      float rho = 0.5f * (ra + rb);
      float E[NR], dE[NR];
#pragma unroll
      for (int q = 0; q < NR; q++) {
        float d = rho - P.Rs_a[q];
        float g = f_ex2(nel2 * d * d);
        E[q] = g;
        dE[q] = -2.0f * P.eta * d * g;
      }
      int p = P.pair_index[ae[j]][ae[k]];
      const float4* Ap = reinterpret_cast<const float4*>(Aa + p * FF_ASTR);
      float W = 0.f, Wt = 0.f, Wr = 0.f;
#pragma unroll
      for (int a = 0; a < NA; a++) {
        float ca = P.cosA[a], sa = P.sinA[a];
        float base = fmaxf(1.0f + c * ca + s * sa, 0.f);
        float T, dT;   // T = pref*base^zeta ; dT = dT/dtheta = -zeta*pref*base^(zeta-1) * sin(theta-theta_a)
        float sind = s * ca - c * sa;
        if (P.zeta_is8) {
          float b2 = base * base, b4 = b2 * b2;
          float b7 = b4 * b2 * base;
          T = P.zeta_pref * b7 * base;
          dT = -8.0f * P.zeta_pref * b7 * sind;
        } else {
          float bm = powf(base, P.zeta - 1.0f);
          T = P.zeta_pref * bm * base;
          dT = -P.zeta * P.zeta_pref * bm * sind;
        }
        float4 A0 = Ap[2 * a], A1 = Ap[2 * a + 1];
        float ua = A0.x * E[0];
        ua = fmaf(A0.y, E[1], ua); ua = fmaf(A0.z, E[2], ua); ua = fmaf(A0.w, E[3], ua);
        ua = fmaf(A1.x, E[4], ua); ua = fmaf(A1.y, E[5], ua); ua = fmaf(A1.z, E[6], ua); ua = fmaf(A1.w, E[7], ua);
        float va = A0.x * dE[0];
        va = fmaf(A0.y, dE[1], va); va = fmaf(A0.z, dE[2], va); va = fmaf(A0.w, dE[3], va);
        va = fmaf(A1.x, dE[4], va); va = fmaf(A1.y, dE[5], va); va = fmaf(A1.z, dE[6], va); va = fmaf(A1.w, dE[7], va);
        W = fmaf(T, ua, W);
        Wt = fmaf(dT, ua, Wt);
        Wr = fmaf(T, va, Wr);
      }
      // V = W fa fb ; gradients w.r.t. a = x_j - x_i and b = x_k - x_i
      float ff = fa * fb;
      float dVdt = Wt * ff;
      float dVda_r = 0.5f * Wr * ff + W * dfa * fb;   // along a_hat
      float dVdb_r = 0.5f * Wr * ff + W * fa * dfb;   // along b_hat
      float is = (s > 1e-6f) ? f_rcp(s) : 0.f;
      // dtheta/da = -(b_hat - c a_hat)/(|a| s) ; dtheta/db = -(a_hat - c b_hat)/(|b| s)
      float ta = -dVdt * is * ira, tb = -dVdt * is * irb;
      float gax = ta * (ubx - c * uax) + dVda_r * uax;
      float gay = ta * (uby - c * uay) + dVda_r * uay;
      float gaz = ta * (ubz - c * uaz) + dVda_r * uaz;
      float gbx = tb * (uax - c * ubx) + dVdb_r * ubx;
      float gby = tb * (uay - c * uby) + dVdb_r * uby;
      float gbz = tb * (uaz - c * ubz) + dVdb_r * ubz;
      gix -= gax + gbx; giy -= gay + gby; giz -= gaz + gbz;
      GA[lane] = make_float4(gax, gay, gaz, 0.f);
      GB[lane] = make_float4(gbx, gby, gbz, 0.f);
    }
    // k-rows present in this round (t is monotone in k)
    const int kfirst = __shfl_sync(FULL, k, 0), klast = __shfl_sync(FULL, k, cnt - 1);
    __syncwarp();
    auto gather = [&](int n, float& fx, float& fy, float& fz) {
      // as the second member (k = n): the run t = n(n-1)/2 + j, j < n
      int base_n = n * (n - 1) / 2;
      int lo = max(base_n, t0), hi = min(base_n + n, t0 + cnt);
      for (int tt = lo; tt < hi; tt++) {
        float4 g = GB[tt - t0];
        fx += g.x; fy += g.y; fz += g.z;
      }
      // as the first member (j = n): one entry in every row k > n
      for (int kk = max(n + 1, kfirst); kk <= klast; kk++) {
        int tt = kk * (kk - 1) / 2 + n;
        if (tt >= t0 && tt < t0 + cnt) {
          float4 g = GA[tt - t0];
          fx += g.x; fy += g.y; fz += g.z;
        }
      }
    };
    if (lane < nang) gather(lane, f0x, f0y, f0z);
    if (lane + 32 < nang) gather(lane + 32, f1x, f1y, f1z);
    __syncwarp();
  }
  if (lane < nang) {
    int dst = aslot[lane];
    if (dst >= 0) {
      atomicAdd(F + 3 * (int64_t)dst, f0x);
      atomicAdd(F + 3 * (int64_t)dst + 1, f0y);
      atomicAdd(F + 3 * (int64_t)dst + 2, f0z);
    }
  }
  if (lane + 32 < nang) {
    int dst = aslot[lane + 32];
    if (dst >= 0) {
      atomicAdd(F + 3 * (int64_t)dst, f1x);
      atomicAdd(F + 3 * (int64_t)dst + 1, f1y);
      atomicAdd(F + 3 * (int64_t)dst + 2, f1z);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    gix += __shfl_xor_sync(FULL, gix, o);
    giy += __shfl_xor_sync(FULL, giy, o);
    giz += __shfl_xor_sync(FULL, giz, o);
  }
  if (lane == 0) {
    atomicAdd(F + 3 * (int64_t)slot, gix);
    atomicAdd(F + 3 * (int64_t)slot + 1, giy);
    atomicAdd(F + 3 * (int64_t)slot + 2, giz);
  }
}

int tm_launch_force(tm_ctx* c, const SysView& s, int flags) {
  const DevParams& P = c->hp;
  size_t wf = tm_force_smem_floats_per_warp(P);
  size_t smem = wf * 4 * FORCE_WARPS;
  int64_t nreal_slots = s.periodic ? s.nreal : s.nslots;
  int fold = (flags & TM_F_FOLD_IMAGES) ? 1 : 0;
  int blocks = (int)((s.nrows + FORCE_WARPS - 1) / FORCE_WARPS);
  // the ANI-1 default grid takes the fast kernel (TM_FORCE_GENERAL=1 keeps the general one: tests, measurements)
  static const bool general_only = getenv("TM_FORCE_GENERAL") != nullptr;
  if (!general_only && P.nAs == 8 && P.nRs_a == 8 && P.nRs_r == 32 && (P.Dp & 3) == 0) {
    size_t wff = tm_force_fast_smem_floats_per_warp(P);
    size_t smf = wff * 4 * FORCE_WARPS;
    static size_t conf_fast_d[64] = {};
    size_t& conf_fast = conf_fast_d[(c->device >= 0 && c->device < 64) ? c->device : 0];
    if (smf > 48 * 1024 && smf > conf_fast) {
      TM_CUDA(cudaFuncSetAttribute(k_force_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smf));
      conf_fast = smf;
    }
    TM_LAUNCH(k_force_fast, blocks, FORCE_WARPS * 32, smf, c->stream, 
        (const SAtom*)c->b_satom.p, (const int32_t*)c->b_rowsidx.p, (const int32_t*)c->b_rowslot.p, (const int32_t*)c->b_nbcnt.p,
        (const uint32_t*)c->b_nbr.p, s.nrows, P, (const float*)c->b_dG[TM_NET_ENERGY].p, (const float*)c->b_dG[TM_NET_CHARGE].p,
        (const double*)c->b_dedq.p, (const double*)c->b_molacc.p, (const double*)c->b_natom.p, s.maxnatom, nreal_slots, fold, (float*)c->b_F.p,
        (int)wff);
    c->launches++;
    TM_CUDA(cudaGetLastError());
    return TM_OK;
  }
  bool small = (P.nAs <= 8 && P.nRs_a <= 8);
  static size_t conf_small_d[64] = {}, conf_big_d[64] = {};   // per device
  const int dv = (c->device >= 0 && c->device < 64) ? c->device : 0;
  size_t& conf_small = conf_small_d[dv];
  size_t& conf_big = conf_big_d[dv];
  if (small) {
    if (smem > 48 * 1024 && smem > conf_small) {
      TM_CUDA(cudaFuncSetAttribute(k_force<8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      conf_small = smem;
    }
    TM_LAUNCH((k_force<8, 8>), blocks, FORCE_WARPS * 32, smem, c->stream, 
        (const SAtom*)c->b_satom.p, (const int32_t*)c->b_rowsidx.p, (const int32_t*)c->b_rowslot.p, (const int32_t*)c->b_nbcnt.p,
        (const uint32_t*)c->b_nbr.p, s.nrows, P, (const float*)c->b_dG[TM_NET_ENERGY].p, (const float*)c->b_dG[TM_NET_CHARGE].p,
        (const double*)c->b_dedq.p, (const double*)c->b_molacc.p, (const double*)c->b_natom.p, s.maxnatom, nreal_slots, fold, (float*)c->b_F.p,
        (int)wf);
  } else {
    if (smem > 48 * 1024 && smem > conf_big) {
      TM_CUDA(cudaFuncSetAttribute(k_force<TM_MAX_SYM, TM_MAX_SYM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      conf_big = smem;
    }
    TM_LAUNCH((k_force<TM_MAX_SYM, TM_MAX_SYM>), blocks, FORCE_WARPS * 32, smem, c->stream, 
        (const SAtom*)c->b_satom.p, (const int32_t*)c->b_rowsidx.p, (const int32_t*)c->b_rowslot.p, (const int32_t*)c->b_nbcnt.p,
        (const uint32_t*)c->b_nbr.p, s.nrows, P, (const float*)c->b_dG[TM_NET_ENERGY].p, (const float*)c->b_dG[TM_NET_CHARGE].p,
        (const double*)c->b_dedq.p, (const double*)c->b_molacc.p, (const double*)c->b_natom.p, s.maxnatom, nreal_slots, fold, (float*)c->b_F.p,
        (int)wf);
  }
  c->launches++;
  TM_CUDA(cudaGetLastError());
  return TM_OK;
}
