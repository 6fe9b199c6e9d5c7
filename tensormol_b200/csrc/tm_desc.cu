// K2+K3: ANI-1 radial and angular symmetry functions with element / element-pair channels.
//
// Computes what TFSymRSet_Linear_WithEle (RawSymFunc.py:1696-1755, Periodic :1812-1863) and
// TFSymASet_Linear_WithEle (RawSymFunc.py:868-962, Periodic :1054-1136) compute, concatenated as
// TFSymSet_Scattered_Linear_WithEle does (RawSymFunc.py:2248): row layout
//   [ radial: e*nRs_r + s | angular: n_ele*nRs_r + p*(nAs*nRs_a) + a*nRs_a + s ]   (theta-major).
// Reference quirks kept: truncated pi in the cutoff (Q3), no fc(r_jk), each unordered {j,k} once,
// prefactor 2^(1-zeta) (Q4).  cos(theta-theta_a) is expanded as cos t cos ta + sin t sin ta with
// sin t = |a x b|/(|a||b|), which removes acos and is well conditioned near collinear triples; the
// reference's clamp to +-(1-1e-16) (Q5) changes the value by < 2e-8 and is below fp32 resolution.
//
// Mapping: one warp per centre row (rows are element-sorted, so the output is the contiguous
// per-element MLP batch).  Neighbour records (32 B) are fetched coalesced from the cell-sorted
// copy; position differences are formed in float64 and rounded once to fp32.
//   radial : lane = Rs channel, neighbours broadcast by warp shuffle, register accumulators per element
//   angular: phase 1 lanes = triples (geometry, nAs angular and nRs_a radial factors into shared tiles),
//            phase 2 lanes = output channels, register accumulators per element-pair channel
// The kernel is instruction-issue bound (ncu r01: 79 % issue-active), not HBM bound, so the template
// parameters exist to cut instructions: NE = number of elements (selects per neighbour), OPLT = outputs
// per lane of one pair channel (2 for the default 8x8 angular grid).  Exponentials use ex2.approx
// (__expf): relative error < 1e-6 on every term that is not itself < 1e-10 of the row scale.
#include "tm_internal.h"
#include <cuda_fp16.h>
#include <cstdlib>

#define FULL 0xffffffffu
#define DESC_WARPS 8
#define RPL 2   // radial channels per lane (num_r_Rs <= 64)

__device__ __forceinline__ void tri_inv(int t, int& j, int& k) {
  // t = k(k-1)/2 + j with 0 <= j < k
  k = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)t)) * 0.5f);
  while (k * (k - 1) / 2 > t) k--;
  while ((k + 1) * k / 2 <= t) k++;
  j = t - k * (k - 1) / 2;
}

__device__ __forceinline__ float pow_zeta(float b, const DevParams& P) {
  if (P.zeta_is8) {
    float b2 = b * b, b4 = b2 * b2;
    return b4 * b4;
  }
  return powf(b, P.zeta);
}

size_t tm_desc_smem_floats_per_warp(const DevParams& P) {
  return 5 * TM_ANG_CAP + TM_ANG_CAP + 32 * (P.nAs + 1) + 32 * (P.nRs_a + 1) + 32;
}

template <int NE, int OPLT>
__global__ void __launch_bounds__(DESC_WARPS * 32)
k_desc(const SAtom* __restrict__ sat, const int32_t* __restrict__ rowsidx, const int32_t* __restrict__ rowslot,
       const int32_t* __restrict__ nbcnt, const uint32_t* __restrict__ nbr, int64_t nrows, const __grid_constant__ DevParams P,
       float* __restrict__ G, __half* __restrict__ Ghi, __half* __restrict__ Glo, int32_t* __restrict__ flags, int wfloats) {
  TM_PDL_PROLOGUE;
  constexpr int NELEP = NE * (NE + 1) / 2;
  extern __shared__ float smem[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t row = (int64_t)blockIdx.x * DESC_WARPS + warp;
  if (row >= nrows) return;
  float* Grow = G + row * P.Dp;
  // the tensor-core MLP path takes the descriptor row as two fp16 planes x = hi + lo/2048 (tm_gemm_tc.cu): written here
  // together with the fp32 row, which saves a separate split pass over G
  auto put = [&](int i, float v) {
    Grow[i] = v;
    if (Ghi) {
      __half h = __float2half_rn(v);
      Ghi[row * P.Dp + i] = h;
      Glo[row * P.Dp + i] = __float2half_rn((v - __half2float(h)) * 2048.0f);
    }
  };
  int slot = rowslot[row];
  if (slot < 0) {
    for (int i = lane; i < P.Dp; i += 32) put(i, 0.f);
    return;
  }
  float* ws = smem + (size_t)warp * wfloats;
  float* ax = ws;
  float* ay = ax + TM_ANG_CAP;
  float* az = ay + TM_ANG_CAP;
  float* ar = az + TM_ANG_CAP;
  float* afc = ar + TM_ANG_CAP;
  int* ae = (int*)(afc + TM_ANG_CAP);
  float* Tt = (float*)(ae + TM_ANG_CAP);
  const int tstr = P.nAs + 1, estr = P.nRs_a + 1;
  float* Et = Tt + 32 * tstr;
  int* chan = (int*)(Et + 32 * estr);

  SAtom ci = sat[rowsidx[row]];
  int b = (int)row * TM_NB_STRIDE, e = b + nbcnt[row];
  float acc[RPL][NE];
  float rs[RPL];
#pragma unroll
  for (int k = 0; k < RPL; k++) {
    rs[k] = (lane + 32 * k < P.nRs_r) ? P.Rs_r[lane + 32 * k] : 0.f;
#pragma unroll
    for (int q = 0; q < NE; q++) acc[k][q] = 0.f;
  }
  const float neg_eta = -P.eta;
  int nang = 0;
  for (int j0 = b; j0 < e; j0 += 32) {
    int j = j0 + lane;
    float dx = 0.f, dy = 0.f, dz = 0.f, r = 1.f, fc = 0.f;
    int ej = 0;
    bool isang = false;
    if (j < e) {
      uint32_t en = nbr[j];
      isang = (en >> 31) != 0;
      SAtom a = sat[en & 0x7fffffffu];
      dx = (float)(a.x - ci.x);
      dy = (float)(a.y - ci.y);
      dz = (float)(a.z - ci.z);
      r = sqrtf(dx * dx + dy * dy + dz * dz);
      ej = a.e;
      fc = 0.5f * (__cosf(P.pi_over_rRc * r) + 1.0f);
      if (P.skin_on) {   // the rows reach out to cutoff + skin (tm_set_skin): the cutoffs are applied here
        if (!(r < P.r_Rc)) fc = 0.f;
        isang = isang && (r < P.a_Rc);
      }
    }
    unsigned mk = __ballot_sync(FULL, isang);
    if (isang) {
      int pos = nang + __popc(mk & ((1u << lane) - 1));
      if (pos < TM_ANG_CAP) {
        ax[pos] = dx; ay[pos] = dy; az[pos] = dz; ar[pos] = r;
        afc[pos] = 0.5f * (__cosf(P.pi_over_aRc * r) + 1.0f);
        ae[pos] = ej;
      } else {
        atomicOr(flags, 4);
      }
    }
    nang += __popc(mk);
    int cnt = min(32, e - j0);
    for (int t = 0; t < cnt; t++) {
      float rr = __shfl_sync(FULL, r, t), ff = __shfl_sync(FULL, fc, t);
      int ee = __shfl_sync(FULL, ej, t);
#pragma unroll
      for (int k = 0; k < RPL; k++) {
        if (k == 0 || P.nRs_r > 32) {
          float d = rr - rs[k];
          float v = __expf(neg_eta * d * d) * ff;
#pragma unroll
          for (int q = 0; q < NE; q++) acc[k][q] += (ee == q) ? v : 0.f;
        }
      }
    }
  }
  nang = min(nang, TM_ANG_CAP);
  __syncwarp();

  // write the radial block now (frees registers for the angular accumulators)
#pragma unroll
  for (int k = 0; k < RPL; k++) {
    int s = lane + 32 * k;
    if (s < P.nRs_r) {
#pragma unroll
      for (int q = 0; q < NE; q++)
        if (q < P.n_ele) put(q * P.nRs_r + s, acc[k][q]);
    }
  }

  // per-lane output coordinates for phase 2
  int oa[OPLT], os[OPLT];
  float ga[NELEP][OPLT];
#pragma unroll
  for (int k = 0; k < OPLT; k++) {
    int idx = lane + 32 * k;
    int a_ = idx / P.nRs_a;
    oa[k] = (idx < P.nsym) ? a_ : 0;
    os[k] = (idx < P.nsym) ? idx - a_ * P.nRs_a : 0;
#pragma unroll
    for (int q = 0; q < NELEP; q++) ga[q][k] = 0.f;
  }
  int ntrip = nang * (nang - 1) / 2;
  for (int t0 = 0; t0 < ntrip; t0 += 32) {
    int t = t0 + lane;
    if (t < ntrip) {
      int j, k;
      tri_inv(t, j, k);
      float ajx = ax[j], ajy = ay[j], ajz = az[j], akx = ax[k], aky = ay[k], akz = az[k];
      float ra = ar[j], rb = ar[k];
      float inv = __frcp_rn(ra * rb);
      float c = (ajx * akx + ajy * aky + ajz * akz) * inv;
      float nx = ajy * akz - ajz * aky, ny = ajz * akx - ajx * akz, nz = ajx * aky - ajy * akx;
      float s = sqrtf(nx * nx + ny * ny + nz * nz) * inv;
      c = fminf(1.0f, fmaxf(-1.0f, c));
      float f = afc[j] * afc[k];
      float rho = 0.5f * (ra + rb);
      for (int a = 0; a < P.nAs; a++) {
        float base = fmaxf(1.0f + c * P.cosA[a] + s * P.sinA[a], 0.f);
        Tt[lane * tstr + a] = P.zeta_pref * pow_zeta(base, P);
      }
      for (int q = 0; q < P.nRs_a; q++) {
        float d = rho - P.Rs_a[q];
        Et[lane * estr + q] = __expf(neg_eta * d * d) * f;
      }
      chan[lane] = P.pair_index[ae[j]][ae[k]];
    }
    __syncwarp();
    int cnt = min(32, ntrip - t0);
    for (int tt = 0; tt < cnt; tt++) {
      int p = chan[tt];
      const float* Tr = Tt + tt * tstr;
      const float* Er = Et + tt * estr;
      float v[OPLT];
#pragma unroll
      for (int k = 0; k < OPLT; k++) v[k] = Tr[oa[k]] * Er[os[k]];
#pragma unroll
      for (int q = 0; q < NELEP; q++) {
        if (p == q) {
#pragma unroll
          for (int k = 0; k < OPLT; k++) ga[q][k] += v[k];
        }
      }
    }
    __syncwarp();
  }

  // angular block, zero padding
  int off = P.n_ele * P.nRs_r;
#pragma unroll
  for (int q = 0; q < NELEP; q++) {
    if (q < P.n_elep) {
#pragma unroll
      for (int k = 0; k < OPLT; k++) {
        int idx = lane + 32 * k;
        if (idx < P.nsym) put(off + q * P.nsym + idx, ga[q][k]);
      }
    }
  }
  for (int i = P.D + lane; i < P.Dp; i += 32) put(i, 0.f);
}

// ---- fast path: the ANI-1 default grid (8 angular x 8 radial functions per pair channel, <= 32 radial functions) ----
// Same mapping and arithmetic as k_desc, with the instruction count per neighbour / triple cut down (k_desc executes
// 3.6k warp instructions per water centre, 2/3 of them in the two accumulation loops):
//   radial : each 32-neighbour chunk is staged in shared memory as r[t] and fc[t] * [e_t == q] (q < 4), so that one
//            neighbour costs two broadcast loads, d, d^2, ex2 and NE FFMAs (no shuffles, no selects);
//   angular: lane l owns the adjacent outputs (a = l / 4, s = 2 (l % 4) + {0, 1}) of a pair channel, so that one triple
//            costs the channel id, T[a] (32-bit) and E[s], E[s+1] (64-bit) loads and two FFMAs under a uniform branch
//            on the channel; the row is written with 64-bit / half2 stores.
__device__ __forceinline__ float ex2_fast(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

template <int NE>
__global__ void __launch_bounds__(DESC_WARPS * 32)
k_desc_fast(const SAtom* __restrict__ sat, const int32_t* __restrict__ rowsidx, const int32_t* __restrict__ rowslot,
            const int32_t* __restrict__ nbcnt, const uint32_t* __restrict__ nbr, int64_t nrows, const __grid_constant__ DevParams P,
            float* __restrict__ G, __half* __restrict__ Ghi, __half* __restrict__ Glo, int32_t* __restrict__ flags, int wfloats) {
  TM_PDL_PROLOGUE;
  constexpr int NELEP = NE * (NE + 1) / 2;
  constexpr int NA = 8, NR = 8, NSYM = NA * NR;
  extern __shared__ float smem[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t row = (int64_t)blockIdx.x * DESC_WARPS + warp;
  if (row >= nrows) return;
  float* Grow = G + row * P.Dp;
  auto put = [&](int i, float v) {
    Grow[i] = v;
    if (Ghi) {
      __half h = __float2half_rn(v);
      Ghi[row * P.Dp + i] = h;
      Glo[row * P.Dp + i] = __float2half_rn((v - __half2float(h)) * 2048.0f);
    }
  };
  auto put2 = [&](int i, float v0, float v1) {   // i even
    *reinterpret_cast<float2*>(Grow + i) = make_float2(v0, v1);
    if (Ghi) {
      __half2 h = __floats2half2_rn(v0, v1);
      float2 hf = __half22float2(h);
      *reinterpret_cast<__half2*>(Ghi + row * P.Dp + i) = h;
      *reinterpret_cast<__half2*>(Glo + row * P.Dp + i) = __floats2half2_rn((v0 - hf.x) * 2048.0f, (v1 - hf.y) * 2048.0f);
    }
  };
  int slot = rowslot[row];
  if (slot < 0) {
    for (int i = lane; i < P.Dp; i += 32) put(i, 0.f);
    return;
  }
  float* ws = smem + (size_t)warp * wfloats;
  float* ax = ws;
  float* ay = ax + TM_ANG_CAP;
  float* az = ay + TM_ANG_CAP;
  float* ar = az + TM_ANG_CAP;
  float* afc = ar + TM_ANG_CAP;
  int* ae = (int*)(afc + TM_ANG_CAP);
  float* U = (float*)(ae + TM_ANG_CAP);   // radial staging, then the angular factor tiles
  float* rr = U;                           // [32]
  float4* ffq = (float4*)(U + 32);         // [32]: fc * [e == q]
  float* Tt = U;                           // [32][NA]
  float* Et = U + 32 * NA;                 // [32][NR]
  int* chan = (int*)(Et + 32 * NR);        // [32]

  SAtom ci = sat[rowsidx[row]];
  int b = (int)row * TM_NB_STRIDE, e = b + nbcnt[row];
  float acc[NE];
#pragma unroll
  for (int q = 0; q < NE; q++) acc[q] = 0.f;
  const float rs0 = (lane < P.nRs_r) ? P.Rs_r[lane] : 0.f;
  const float c2 = -P.eta * 1.4426950408889634f;   // exp(-eta x) = 2^(c2 x)
  int nang = 0;
  for (int j0 = b; j0 < e; j0 += 32) {
    int j = j0 + lane;
    float dx = 0.f, dy = 0.f, dz = 0.f, r = 1.f, fc = 0.f;
    int ej = 0;
    bool isang = false;
    if (j < e) {
      uint32_t en = nbr[j];
      isang = (en >> 31) != 0;
      SAtom a = sat[en & 0x7fffffffu];
      dx = (float)(a.x - ci.x);
      dy = (float)(a.y - ci.y);
      dz = (float)(a.z - ci.z);
      r = sqrtf(dx * dx + dy * dy + dz * dz);
      ej = a.e;
      fc = 0.5f * (__cosf(P.pi_over_rRc * r) + 1.0f);
      if (P.skin_on) {   // the rows reach out to cutoff + skin (tm_set_skin): the cutoffs are applied here
        if (!(r < P.r_Rc)) fc = 0.f;
        isang = isang && (r < P.a_Rc);
      }
    }
    rr[lane] = r;
    ffq[lane] = make_float4(ej == 0 ? fc : 0.f, ej == 1 ? fc : 0.f, ej == 2 ? fc : 0.f, ej == 3 ? fc : 0.f);
    unsigned mk = __ballot_sync(FULL, isang);
    if (isang) {
      int pos = nang + __popc(mk & ((1u << lane) - 1));
      if (pos < TM_ANG_CAP) {
        ax[pos] = dx; ay[pos] = dy; az[pos] = dz; ar[pos] = r;
        afc[pos] = 0.5f * (__cosf(P.pi_over_aRc * r) + 1.0f);
        ae[pos] = ej;
      } else {
        atomicOr(flags, 4);
      }
    }
    nang += __popc(mk);
    __syncwarp();
    int cnt = min(32, e - j0);
#pragma unroll 4
    for (int t = 0; t < cnt; t++) {
      float d = rr[t] - rs0;
      float4 f = ffq[t];
      float v = ex2_fast(c2 * d * d);
      acc[0] = fmaf(v, f.x, acc[0]);
      if (NE > 1) acc[1] = fmaf(v, f.y, acc[1]);
      if (NE > 2) acc[2] = fmaf(v, f.z, acc[2]);
      if (NE > 3) acc[3] = fmaf(v, f.w, acc[3]);
    }
    __syncwarp();
  }
  nang = min(nang, TM_ANG_CAP);

  if (lane < P.nRs_r) {
#pragma unroll
    for (int q = 0; q < NE; q++)
      if (q < P.n_ele) put(q * P.nRs_r + lane, acc[q]);
  }

  float ga[NELEP][2];
#pragma unroll
  for (int q = 0; q < NELEP; q++) ga[q][0] = ga[q][1] = 0.f;
  const int la = lane >> 2, ls = 2 * (lane & 3);
  int ntrip = nang * (nang - 1) / 2;
  for (int t0 = 0; t0 < ntrip; t0 += 32) {
    int t = t0 + lane;
    if (t < ntrip) {
      int j, k;
      tri_inv(t, j, k);
      float ajx = ax[j], ajy = ay[j], ajz = az[j], akx = ax[k], aky = ay[k], akz = az[k];
      float ra = ar[j], rb = ar[k];
      float inv = __frcp_rn(ra * rb);
      float c = (ajx * akx + ajy * aky + ajz * akz) * inv;
      float nx = ajy * akz - ajz * aky, ny = ajz * akx - ajx * akz, nz = ajx * aky - ajy * akx;
      float s = sqrtf(nx * nx + ny * ny + nz * nz) * inv;
      c = fminf(1.0f, fmaxf(-1.0f, c));
      float f = afc[j] * afc[k];
      float rho = 0.5f * (ra + rb);
      float tv[NA], ev[NR];
#pragma unroll
      for (int a = 0; a < NA; a++) {
        float base = fmaxf(1.0f + c * P.cosA[a] + s * P.sinA[a], 0.f);
        tv[a] = P.zeta_pref * pow_zeta(base, P);
      }
#pragma unroll
      for (int q = 0; q < NR; q++) {
        float d = rho - P.Rs_a[q];
        ev[q] = ex2_fast(c2 * d * d) * f;
      }
      float4* Tw = reinterpret_cast<float4*>(Tt + lane * NA);
      float4* Ew = reinterpret_cast<float4*>(Et + lane * NR);
      Tw[0] = make_float4(tv[0], tv[1], tv[2], tv[3]);
      Tw[1] = make_float4(tv[4], tv[5], tv[6], tv[7]);
      Ew[0] = make_float4(ev[0], ev[1], ev[2], ev[3]);
      Ew[1] = make_float4(ev[4], ev[5], ev[6], ev[7]);
      chan[lane] = P.pair_index[ae[j]][ae[k]];
    }
    __syncwarp();
    int cnt = min(32, ntrip - t0);
#pragma unroll 4
    for (int tt = 0; tt < cnt; tt++) {
      int p = chan[tt];
      float ta = Tt[tt * NA + la];
      float2 e2 = *reinterpret_cast<const float2*>(Et + tt * NR + ls);
#pragma unroll
      for (int q = 0; q < NELEP; q++) {
        if (p == q) {
          ga[q][0] = fmaf(ta, e2.x, ga[q][0]);
          ga[q][1] = fmaf(ta, e2.y, ga[q][1]);
        }
      }
    }
    __syncwarp();
  }

  // angular block (lane owns outputs 2 lane, 2 lane + 1 of every pair channel), zero padding
  int off = P.n_ele * P.nRs_r;
#pragma unroll
  for (int q = 0; q < NELEP; q++)
    if (q < P.n_elep) put2(off + q * NSYM + 2 * lane, ga[q][0], ga[q][1]);
  for (int i = P.D + lane; i < P.Dp; i += 32) put(i, 0.f);
}

template <int NE>
static int launch_desc_fast(tm_ctx* c, const SysView& s) {
  const DevParams& P = c->hp;
  size_t wf = tm_desc_smem_floats_per_warp(P);
  size_t smem = wf * 4 * DESC_WARPS;
  static size_t configured[64] = {};            // per device
  const int dv = (c->device >= 0 && c->device < 64) ? c->device : 0;
  if (smem > 48 * 1024 && smem > configured[dv]) {
    TM_CUDA(cudaFuncSetAttribute(k_desc_fast<NE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dv] = smem;
  }
  int blocks = (int)((s.nrows + DESC_WARPS - 1) / DESC_WARPS);
  const bool split = c->gemm_mode != TM_GEMM_FP32;
  TM_LAUNCH(k_desc_fast<NE>, blocks, DESC_WARPS * 32, smem, c->stream, (const SAtom*)c->b_satom.p, (const int32_t*)c->b_rowsidx.p,
                                                               (const int32_t*)c->b_rowslot.p, (const int32_t*)c->b_nbcnt.p,
                                                               (const uint32_t*)c->b_nbr.p, s.nrows, P, (float*)c->b_G.p,
                                                               split ? (__half*)c->b_Gs.p : nullptr,
                                                               split ? (__half*)c->b_Gs.p + (size_t)s.nrows * P.Dp : nullptr,
                                                               (int32_t*)c->b_flags.p, (int)wf);
  c->launches++;
  TM_CUDA(cudaGetLastError());
  return TM_OK;
}

template <int NE, int OPLT>
static int launch_desc(tm_ctx* c, const SysView& s) {
  const DevParams& P = c->hp;
  size_t wf = tm_desc_smem_floats_per_warp(P);
  size_t smem = wf * 4 * DESC_WARPS;
  static size_t configured[64] = {};            // per device
  const int dv = (c->device >= 0 && c->device < 64) ? c->device : 0;
  if (smem > 48 * 1024 && smem > configured[dv]) {
    TM_CUDA(cudaFuncSetAttribute(k_desc<NE, OPLT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dv] = smem;
  }
  int blocks = (int)((s.nrows + DESC_WARPS - 1) / DESC_WARPS);
  const bool split = c->gemm_mode != TM_GEMM_FP32;
  TM_LAUNCH((k_desc<NE, OPLT>), blocks, DESC_WARPS * 32, smem, c->stream, (const SAtom*)c->b_satom.p, (const int32_t*)c->b_rowsidx.p,
                                                                (const int32_t*)c->b_rowslot.p, (const int32_t*)c->b_nbcnt.p,
                                                                (const uint32_t*)c->b_nbr.p, s.nrows, P, (float*)c->b_G.p,
                                                                split ? (__half*)c->b_Gs.p : nullptr,
                                                                split ? (__half*)c->b_Gs.p + (size_t)s.nrows * P.Dp : nullptr,
                                                                (int32_t*)c->b_flags.p, (int)wf);
  c->launches++;
  TM_CUDA(cudaGetLastError());
  return TM_OK;
}

int tm_launch_desc(tm_ctx* c, const SysView& s) {
  int rc;
  const DevParams& P = c->hp;
  if ((rc = tm_buf(c, c->b_G, (size_t)s.nrows * P.Dp * 4))) return rc;
  if (c->gemm_mode != TM_GEMM_FP32)
    if ((rc = tm_buf(c, c->b_Gs, (size_t)2 * s.nrows * P.Dp * 2))) return rc;   // [hi | lo] fp16 planes
  // the ANI-1 default grid takes the fast kernel (TM_DESC_GENERAL=1 keeps the general one: tests, measurements);
  // its row layout needs an even radial block so that the 64-bit stores of the angular block stay aligned
  static const bool general_only = getenv("TM_DESC_GENERAL") != nullptr;
  if (!general_only && P.nAs == 8 && P.nRs_a == 8 && P.nRs_r <= 32 && P.n_ele <= 4 && ((P.n_ele * P.nRs_r) & 1) == 0 && (P.Dp & 1) == 0) {
    switch (P.n_ele) {
      case 1: return launch_desc_fast<1>(c, s);
      case 2: return launch_desc_fast<2>(c, s);
      case 3: return launch_desc_fast<3>(c, s);
      default: return launch_desc_fast<4>(c, s);
    }
  }
  bool small = P.nsym <= 64;
  if (small && P.n_ele == 1) return launch_desc<1, 2>(c, s);
  if (small && P.n_ele == 2) return launch_desc<2, 2>(c, s);
  if (small && P.n_ele == 3) return launch_desc<3, 2>(c, s);
  if (small && P.n_ele == 4) return launch_desc<4, 2>(c, s);
  if (P.n_ele <= 4) return launch_desc<4, 8>(c, s);
  if (small) return launch_desc<TM_MAX_ELE, 2>(c, s);
  // 8 elements x 256 angular functions would need 36 x 8 register accumulators per lane: not supported
  tm_set_error("more than 4 elements together with more than 64 angular functions per pair channel is not supported");
  return TM_EINVAL;
}
