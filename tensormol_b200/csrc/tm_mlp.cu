// K4: per-element Behler-Parrinello MLPs (charge net + energy net) as grouped GEMMs over the
// element-sorted row batches, forward and backward-data.
//
// Restates energy_inference / dipole_inference (TFMolInstanceDirect.py:5164-5285; periodic
// :5774-5898): per element  D -> H1 -> ... -> Hn -> 1, y = a(xW+b), last layer linear, activation
// sigmoid_with_param = log(1+exp(alpha x))/alpha (Util.py:200-201) evaluated as a stable softplus.
// The backward pass replaces tf.gradients through the nets: delta_L = w_out * a'(h_L),
// delta_{l-1} = (delta_l W_l^T) * a'(h_{l-1}), dy/dG = delta_0 W_0^T, with a' recovered from the stored
// activation (softplus: a'(z) = 1 - exp(-alpha h)).
//
// A "group" is one (net, element): rows [rowmeta[2e], +rowmeta[2e+1]) of the row space (padded to
// TM_ROW_TILE), so tiles never straddle elements and row counts stay on the device.
#include "tm_internal.h"
#include <cuda_fp16.h>
#include <algorithm>

#define FULL 0xffffffffu

struct GemmGroupTbl {
  GemmGroup g[2 * TM_MAX_ELE];
};

__device__ __forceinline__ float act_fwd(float z, int kind, float alpha) {
  switch (kind) {
    case TM_ACT_SIGMOID_WITH_PARAM: {
      float t = alpha * z;
      return (fmaxf(t, 0.f) + log1pf(expf(-fabsf(t)))) / alpha;
    }
    case TM_ACT_RELU: return fmaxf(z, 0.f);
    case TM_ACT_SOFTPLUS: return fmaxf(z, 0.f) + log1pf(expf(-fabsf(z)));
    case TM_ACT_TANH: return tanhf(z);
    case TM_ACT_ELU: return z > 0.f ? z : expm1f(z);
    case TM_ACT_SELU: return z >= 0.f ? TM_SELU_SCALE * z : TM_SELU_SCALE * TM_SELU_ALPHA * expm1f(z);
    default: return 1.0f / (1.0f + expf(-z));
  }
}
// derivative a'(z) expressed through the stored activation h = a(z)
__device__ __forceinline__ float act_bwd_from_h(float h, int kind, float alpha) {
  switch (kind) {
    case TM_ACT_SIGMOID_WITH_PARAM: return -expm1f(-alpha * h);
    case TM_ACT_RELU: return h > 0.f ? 1.f : 0.f;
    case TM_ACT_SOFTPLUS: return -expm1f(-h);
    case TM_ACT_TANH: return 1.0f - h * h;
    case TM_ACT_ELU: return h > 0.f ? 1.f : h + 1.0f;
    case TM_ACT_SELU: return h >= 0.f ? TM_SELU_SCALE : h + TM_SELU_SCALE * TM_SELU_ALPHA;
    default: return h * (1.0f - h);
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 FFMA GEMM: 128x128x16 tiles, 256 threads, 8x8 register tile per thread, register-staged
// double buffering.  C = epi(A[rows,K] * B[K,N]).  K % 16 == 0, N % 128 == 0, rows padded to 128.
// ------------------------------------------------------------------------------------------------
#define BM 128
#define BN 128
#define BK 16
#define APAD 4

__global__ void __launch_bounds__(256)
k_gemm_simt(const __grid_constant__ GemmGroupTbl tbl, const int32_t* __restrict__ rowmeta, int epilogue, int act_kind, float act_alpha) {
  TM_PDL_PROLOGUE;
  const GemmGroup& G = tbl.g[blockIdx.z];
  int rows_e = rowmeta[2 * G.ele + 1];
  int rt = blockIdx.x;
  if (rt * BM >= rows_e) return;
  int nt = blockIdx.y;
  if (nt * BN >= G.N) return;
  int64_t row0 = (int64_t)rowmeta[2 * G.ele] + (int64_t)rt * BM;
  const float* A = (const float*)G.A + row0 * G.lda;
  const float* B = (const float*)G.B + nt * BN;
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN];
  int tid = threadIdx.x;
  int tx = tid & 15, ty = tid >> 4;
  // global->register staging coordinates
  int a_r = tid >> 2, a_k = (tid & 3) * 4;   // rows a_r and a_r+64, 4 consecutive k
  int b_k = tid >> 5, b_n = (tid & 31) * 4;  // k rows b_k and b_k+8, 4 consecutive n
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
  float4 ra0, ra1, rb0, rb1;
  int nk = G.K / BK;
  auto gload = [&](int kt) {
    const float* Ap = A + (int64_t)a_r * G.lda + kt * BK + a_k;
    ra0 = *reinterpret_cast<const float4*>(Ap);
    ra1 = *reinterpret_cast<const float4*>(Ap + (int64_t)64 * G.lda);
    const float* Bp = B + (int64_t)(kt * BK + b_k) * G.ldb + b_n;
    rb0 = *reinterpret_cast<const float4*>(Bp);
    rb1 = *reinterpret_cast<const float4*>(Bp + (int64_t)8 * G.ldb);
  };
  auto sstore = [&](int buf) {
    As[buf][a_k + 0][a_r] = ra0.x; As[buf][a_k + 1][a_r] = ra0.y; As[buf][a_k + 2][a_r] = ra0.z; As[buf][a_k + 3][a_r] = ra0.w;
    As[buf][a_k + 0][a_r + 64] = ra1.x; As[buf][a_k + 1][a_r + 64] = ra1.y; As[buf][a_k + 2][a_r + 64] = ra1.z; As[buf][a_k + 3][a_r + 64] = ra1.w;
    *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n]) = rb0;
    *reinterpret_cast<float4*>(&Bs[buf][b_k + 8][b_n]) = rb1;
  };
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; kt++) {
    int buf = kt & 1;
    if (kt + 1 < nk) gload(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  // epilogue
#pragma unroll
  for (int i = 0; i < 8; i++) {
    int r = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4);
    int64_t grow = row0 + r;
#pragma unroll
    for (int half = 0; half < 2; half++) {
      int cn = nt * BN + half * 64 + tx * 4;
      float v[4] = {acc[i][half * 4 + 0], acc[i][half * 4 + 1], acc[i][half * 4 + 2], acc[i][half * 4 + 3]};
      if (epilogue == TM_EPI_ACT) {
        float4 bb = *reinterpret_cast<const float4*>(G.bias + cn);
        v[0] = act_fwd(v[0] + bb.x, act_kind, act_alpha);
        v[1] = act_fwd(v[1] + bb.y, act_kind, act_alpha);
        v[2] = act_fwd(v[2] + bb.z, act_kind, act_alpha);
        v[3] = act_fwd(v[3] + bb.w, act_kind, act_alpha);
      } else if (epilogue == TM_EPI_DACT) {
        float4 hh = *reinterpret_cast<const float4*>((const float*)G.Hmul + grow * G.ldc + cn);
        v[0] *= act_bwd_from_h(hh.x, act_kind, act_alpha);
        v[1] *= act_bwd_from_h(hh.y, act_kind, act_alpha);
        v[2] *= act_bwd_from_h(hh.z, act_kind, act_alpha);
        v[3] *= act_bwd_from_h(hh.w, act_kind, act_alpha);
      }
      *reinterpret_cast<float4*>((float*)G.C + grow * G.ldc + cn) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

int tm_gemm_tc_launch(tm_ctx* c, const GemmGroup* groups, int ngroups, const int* rowmeta_dev, int max_row_tiles, int64_t expect_rows, int epilogue);

int tm_launch_gemm(tm_ctx* c, const GemmGroup* groups, int ngroups, const int* rowmeta_dev, int max_row_tiles, int64_t expect_rows, int epilogue) {
  if (c->gemm_mode != TM_GEMM_FP32) return tm_gemm_tc_launch(c, groups, ngroups, rowmeta_dev, max_row_tiles, expect_rows, epilogue);
  GemmGroupTbl tbl;
  int maxN = 0;
  for (int i = 0; i < ngroups; i++) {
    tbl.g[i] = groups[i];
    if (groups[i].N > maxN) maxN = groups[i].N;
    if (groups[i].K % BK || groups[i].N % BN) {
      tm_set_error("gemm dims not padded: K=%d N=%d", groups[i].K, groups[i].N);
      return TM_EINVAL;
    }
  }
  dim3 grid((unsigned)max_row_tiles, (unsigned)(maxN / BN), (unsigned)ngroups);
  TM_LAUNCH(k_gemm_simt, grid, 256, 0, c->stream, tbl, rowmeta_dev, epilogue, c->hp.activation, c->hp.act_alpha);
  c->launches++;
  TM_CUDA(cudaGetLastError());
  return TM_OK;
}


// ------------------------------------------------------------------------------------------------
// output layer (H_last -> 1) and its backward seed (fp32 mode), and the reduction of the fused output layer's partial sums (tensor-core mode)
// ------------------------------------------------------------------------------------------------
struct OutTbl {
  const float* w[2][TM_MAX_ELE];
  float b[2][TM_MAX_ELE];
  const float* h[2];      // last hidden activation [nrows][ld]
  float* y[2];            // [nrows]
  float* delta[2];        // [nrows][ld]
  int ld, H;
};

__device__ __forceinline__ int row_element(const int32_t* rowmeta, int64_t row, int n_ele) {
  for (int e = 0; e < n_ele; e++) {
    int b = rowmeta[2 * e], n = rowmeta[2 * e + 1];
    if (row >= b && row < b + n) return e;
  }
  return -1;
}

// fp32 mode only (the tensor-core path fuses this into the last hidden layer's GEMM epilogue, TM_EPI_ACT_OUT):
// one warp per (row, net): y = h . w + b ; delta = w * a'(h)
__global__ void k_out_layer(const __grid_constant__ OutTbl T, const int32_t* __restrict__ rowmeta, int64_t nrows, int n_ele, int act_kind,
                            float act_alpha) {
  TM_PDL_PROLOGUE;
  int64_t wid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  int net = (int)(wid & 1);
  int64_t row = wid >> 1;
  if (row >= nrows) return;
  int e = row_element(rowmeta, row, n_ele);
  if (e < 0) { if (lane == 0) T.y[net][row] = 0.f; return; }
  const float* w = T.w[net][e];
  float s = 0.f;
  for (int i = lane * 4; i < T.ld; i += 128) {
    float4 hv = *reinterpret_cast<const float4*>(T.h[net] + row * T.ld + i);
    float4 wv = *reinterpret_cast<const float4*>(w + i);
    s += hv.x * wv.x + hv.y * wv.y + hv.z * wv.z + hv.w * wv.w;
    *reinterpret_cast<float4*>(T.delta[net] + row * T.ld + i) =
        make_float4(wv.x * act_bwd_from_h(hv.x, act_kind, act_alpha), wv.y * act_bwd_from_h(hv.y, act_kind, act_alpha),
                    wv.z * act_bwd_from_h(hv.z, act_kind, act_alpha), wv.w * act_bwd_from_h(hv.w, act_kind, act_alpha));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
  if (lane == 0) T.y[net][row] = s + T.b[net][e];
}

// y[net][row] = b_out + sum of the np partial dot products left by the TM_EPI_ACT_OUT epilogue (fixed order: deterministic).
// The same pass hands the results on (these were three more launches): q_raw into slot order (qraw_slot != nullptr;
// TFMolInstanceDirect.py:5274), its per-molecule sum for the neutralisation (molacc[16m+4], sum_q != 0) and the
// per-molecule sum of the atomic energies (molacc[16m+1] = Ebp), one atomic per block and quantity when the block's
// rows belong to one molecule.
struct YTbl {
  float b[2][TM_MAX_ELE];
  const float* part[2];
  float* y[2];
};
__global__ void __launch_bounds__(256)
k_y_reduce(const __grid_constant__ YTbl T, const int32_t* __restrict__ rowmeta, int64_t nrows, int n_ele, int np,
           const int32_t* __restrict__ rowslot, int64_t maxnatom, double* __restrict__ qraw_slot, double* __restrict__ molacc, int sum_q) {
  TM_PDL_PROLOGUE;
  __shared__ int s_m[8];
  __shared__ double s_q[8], s_e[8];
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool live = t < 2 * nrows;
  int net = live ? (int)(t / nrows) : 0;
  int64_t row = live ? t - (int64_t)net * nrows : 0;
  int e = live ? row_element(rowmeta, row, n_ele) : -1;
  float s = 0.f;
  if (e >= 0) {
    s = T.b[net][e];
    for (int p = 0; p < np; p++) s += T.part[net][(int64_t)p * nrows + row];
  }
  if (live) T.y[net][row] = s;
  int slot = (e >= 0) ? rowslot[row] : -1;
  int m = -1;
  double vq = 0.0, ve = 0.0;
  if (slot >= 0) {
    m = (int)(slot / maxnatom);
    if (net == TM_NET_CHARGE) {
      vq = (double)s;
      if (qraw_slot) qraw_slot[slot] = vq;
    } else {
      ve = (double)s;
    }
  }
  // warp, then block aggregation; a warp that straddles molecules falls back to one atomic per thread
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int mmax = m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mmax = max(mmax, __shfl_xor_sync(FULL, mmax, o));
  const bool uniform = __all_sync(FULL, m == mmax || m < 0);
  if (uniform) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      vq += __shfl_xor_sync(FULL, vq, o);
      ve += __shfl_xor_sync(FULL, ve, o);
    }
  } else if (m >= 0) {
    if (net == TM_NET_CHARGE) { if (sum_q) atomicAdd(&molacc[16 * m + 4], vq); }
    else atomicAdd(&molacc[16 * m + 1], ve);
  }
  if (lane == 0) { s_m[w] = uniform ? mmax : -1; s_q[w] = vq; s_e[w] = ve; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int cur = -1;
    double aq = 0.0, ae = 0.0;
    for (int k = 0; k <= 8; k++) {
      int mk = (k < 8) ? s_m[k] : -2;
      if (mk == -1) continue;
      if (mk != cur) {
        if (cur >= 0) {
          if (sum_q) atomicAdd(&molacc[16 * cur + 4], aq);
          atomicAdd(&molacc[16 * cur + 1], ae);
        }
        cur = mk; aq = 0.0; ae = 0.0;
      }
      if (k < 8) { aq += s_q[k]; ae += s_e[k]; }
    }
  }
}

// Buffer planes: every activation / delta buffer holds either one fp32 plane (fp32 mode) or two fp16 planes [hi | lo]
// (tensor-core mode); both need the same bytes.
static int ensure_mlp_bufs(tm_ctx* c, const SysView& s) {
  int rc;
  int nh = c->desc.n_hidden;
  for (int net = 0; net < 2; net++) {
    for (int l = 0; l < nh; l++)
      if ((rc = tm_buf(c, c->b_act[net][l], (size_t)s.nrows * c->Hp[l] * 4))) return rc;
    if ((rc = tm_buf(c, c->b_y[net], (size_t)s.nrows * 4))) return rc;
    if ((rc = tm_buf(c, c->b_dG[net], (size_t)s.nrows * c->hp.Dp * 4))) return rc;
  }
  if ((rc = tm_buf(c, c->b_delta0, (size_t)2 * s.nrows * c->Hmax * 4))) return rc;   // [plane][net][nrows*Hmax] fp16, or [net][nrows*Hmax] fp32
  if ((rc = tm_buf(c, c->b_delta1, (size_t)2 * s.nrows * c->Hmax * 4))) return rc;
  if (c->gemm_mode != TM_GEMM_FP32) {
    if ((rc = tm_buf(c, c->b_Gs, (size_t)2 * s.nrows * c->hp.Dp * 2))) return rc;
    if ((rc = tm_buf(c, c->b_ypart, (size_t)2 * (2 * c->Hmax / 128) * s.nrows * 4))) return rc;   // [net][2*N/128][nrows]
  }
  return TM_OK;
}

// delta for hidden layer l (net, plane): buffers alternate with l
static void* delta_ptr(tm_ctx* c, const SysView& s, int l, int net, int plane) {
  char* base = (char*)((l & 1) ? c->b_delta1.p : c->b_delta0.p);
  size_t esz = (c->gemm_mode != TM_GEMM_FP32) ? 2 : 4;
  return base + ((size_t)plane * 2 + net) * s.nrows * c->Hmax * esz;
}

// centres this rank is likely to own (the exact count lives on the device): sizes the GEMM column tiles
static int64_t expect_rows(const SysView& s) {
  return std::max<int64_t>(1, s.slab_world > 1 ? (s.periodic ? s.nreal : s.nslots) / s.slab_world : s.ncent_max);
}

int tm_launch_mlp_forward(tm_ctx* c, const SysView& s) {
  int rc;
  if ((rc = ensure_mlp_bufs(c, s))) return rc;
  const bool tc = c->gemm_mode != TM_GEMM_FP32;
  int nh = c->desc.n_hidden, ne = c->hp.n_ele;
  int max_tiles = (int)((s.ncent_max + TM_ROW_TILE - 1) / TM_ROW_TILE) + 1;
  const int* rowmeta = (const int*)c->b_rowmeta.p;
  // (tensor-core mode: k_desc already wrote the descriptor rows as fp16 hi/lo planes into b_Gs)
  GemmGroup all[TM_MAX_HIDDEN][2 * TM_MAX_ELE];
  int epis[TM_MAX_HIDDEN];
  int ng = 0;
  for (int l = 0; l < nh; l++) {
    GemmGroup* gg = all[l];
    ng = 0;
    for (int net = 0; net < 2; net++)
      for (int e = 0; e < ne; e++) {
        const Layer& L = c->nets[net][e].layers[l];
        GemmGroup& g = gg[ng++];
        g = GemmGroup();
        int lda = (l == 0) ? c->hp.Dp : c->Hp[l - 1];
        size_t aplane = (size_t)s.nrows * lda, cplane = (size_t)s.nrows * L.Np;
        if (tc) {
          const uint16_t* a = (l == 0) ? (const uint16_t*)c->b_Gs.p : (const uint16_t*)c->b_act[net][l - 1].p;
          g.A = a; g.A2 = a + aplane;
          g.B = L.WTs; g.B2 = L.WTs + (size_t)L.Np * L.Kp; g.ldb = L.Kp;
        } else {
          g.A = (l == 0) ? (const float*)c->b_G.p : (const float*)c->b_act[net][l - 1].p;
          g.B = L.W; g.ldb = L.Np;
        }
        g.lda = lda;
        g.bias = L.b; g.Hmul = nullptr;
        g.C = c->b_act[net][l].p; g.C2 = tc ? (void*)((uint16_t*)g.C + cplane) : nullptr; g.ldc = L.Np;
        g.K = L.Kp; g.N = L.Np; g.ele = e; g.rows_alloc = s.nrows;
        if (tc && l == nh - 1) {   // fused output layer: the seed of the backward pass replaces the stored activation
          g.C = delta_ptr(c, s, l, net, 0); g.C2 = delta_ptr(c, s, l, net, 1);
          g.wout = c->nets[net][e].w_out;
          g.ypart = (float*)c->b_ypart.p + (size_t)net * (2 * c->Hmax / 128) * s.nrows;
        }
      }
    epis[l] = (tc && l == nh - 1) ? TM_EPI_ACT_OUT : TM_EPI_ACT;
  }
  // tensor-core mode: all layers in one persistent launch when the fused kernel applies, else layer by layer
  rc = 1;
  if (tc) {
    GemmGroup flat[TM_MAX_HIDDEN * 2 * TM_MAX_ELE];
    for (int l = 0; l < nh; l++)
      for (int i = 0; i < ng; i++) flat[l * ng + i] = all[l][i];
    rc = tm_gemm_tc_launch_multi(c, flat, nh, ng, rowmeta, max_tiles, expect_rows(s), epis, false);
    if (rc < 0) return rc;
    if (rc == 0) tm_trace(c, "forward GEMMs (all layers, one launch)");
  }
  if (rc == 1)
    for (int l = 0; l < nh; l++) {
      if ((rc = tm_launch_gemm(c, all[l], ng, rowmeta, max_tiles, expect_rows(s), epis[l]))) return rc;
      tm_trace(c, l == 0 ? "forward GEMM layer 0" : l == 1 ? "forward GEMM layer 1" : "forward GEMM layer 2+");
    }
  if (tc) {
    YTbl Y;
    for (int net = 0; net < 2; net++) {
      for (int e = 0; e < TM_MAX_ELE; e++) Y.b[net][e] = (e < ne) ? c->nets[net][e].b_out : 0.f;
      Y.part[net] = (const float*)c->b_ypart.p + (size_t)net * (2 * c->Hmax / 128) * s.nrows;
      Y.y[net] = (float*)c->b_y[net].p;
    }
    int np = 2 * c->Hp[nh - 1] / 128;
    // single-device runs: q_raw goes straight into slot order (slab ranks exchange it first, tm_slab_phase_a)
    double* qraw = nullptr;
    if (!s.slab_api) {
      int64_t nq = s.periodic ? s.nreal : s.nslots;           // slots that carry an own charge
      if ((rc = tm_buf(c, c->b_q, (size_t)nq * 8 * 2))) return rc;   // [qraw_slot | q_slot]
      qraw = (double*)c->b_q.p;
      TM_CUDA(cudaMemsetAsync(qraw, 0, (size_t)nq * 8, c->stream));   // slots without a row (padding) stay 0
    }
    TM_LAUNCH(k_y_reduce, (unsigned)((2 * s.nrows + 255) / 256), 256, 0, c->stream, Y, (const int32_t*)c->b_rowmeta.p, s.nrows, ne, np,
                                                                           (const int32_t*)c->b_rowslot.p, s.maxnatom, qraw, (double*)c->b_molacc.p,
                                                                           s.slab_api ? 0 : 1);
    c->launches++;
    c->y_fused = true;
    TM_CUDA(cudaGetLastError());
    return TM_OK;
  }
  c->y_fused = false;
  OutTbl T;
  for (int net = 0; net < 2; net++) {
    for (int e = 0; e < TM_MAX_ELE; e++) {
      T.w[net][e] = (e < ne) ? c->nets[net][e].w_out : nullptr;
      T.b[net][e] = (e < ne) ? c->nets[net][e].b_out : 0.f;
    }
    T.h[net] = (const float*)c->b_act[net][nh - 1].p;
    T.y[net] = (float*)c->b_y[net].p;
    T.delta[net] = (float*)delta_ptr(c, s, nh - 1, net, 0);
  }
  T.ld = c->Hp[nh - 1];
  T.H = c->desc.hidden[nh - 1];
  int64_t nw = s.nrows * 2;
  int blocks = (int)((nw * 32 + 255) / 256);
  TM_LAUNCH(k_out_layer, blocks, 256, 0, c->stream, T, (const int32_t*)c->b_rowmeta.p, s.nrows, ne, c->hp.activation, c->hp.act_alpha);
  c->launches++;
  TM_CUDA(cudaGetLastError());
  return TM_OK;
}

int tm_launch_mlp_backward(tm_ctx* c, const SysView& s) {
  int rc;
  const bool tc = c->gemm_mode != TM_GEMM_FP32;
  int nh = c->desc.n_hidden, ne = c->hp.n_ele;
  int max_tiles = (int)((s.ncent_max + TM_ROW_TILE - 1) / TM_ROW_TILE) + 1;
  const int* rowmeta = (const int*)c->b_rowmeta.p;
  GemmGroup all[TM_MAX_HIDDEN][2 * TM_MAX_ELE];   // in execution order: index 0 = the last hidden layer
  int epis[TM_MAX_HIDDEN];
  int ng = 0;
  for (int l = nh - 1; l >= 0; l--) {
    GemmGroup* gg = all[nh - 1 - l];
    ng = 0;
    for (int net = 0; net < 2; net++)
      for (int e = 0; e < ne; e++) {
        const Layer& L = c->nets[net][e].layers[l];
        GemmGroup& g = gg[ng++];
        g = GemmGroup();
        g.A = delta_ptr(c, s, l, net, 0); g.A2 = delta_ptr(c, s, l, net, 1); g.lda = c->Hp[l];
        if (tc) { g.B = L.Ws; g.B2 = L.Ws + (size_t)L.Kp * L.Np; g.ldb = L.Np; }   // [N'=Kp][K'=Np], K-major
        else { g.B = L.WT; g.ldb = L.Kp; }
        g.bias = nullptr;
        g.K = L.Np; g.N = L.Kp; g.ele = e; g.rows_alloc = s.nrows;
        if (l > 0) {
          g.Hmul = c->b_act[net][l - 1].p;
          g.Hmul2 = (const uint16_t*)g.Hmul + (size_t)s.nrows * c->Hp[l - 1];
          g.C = delta_ptr(c, s, l - 1, net, 0); g.C2 = delta_ptr(c, s, l - 1, net, 1); g.ldc = c->Hp[l - 1];
        } else {
          g.Hmul = nullptr;
          g.C = c->b_dG[net].p; g.ldc = c->hp.Dp;
        }
      }
    epis[nh - 1 - l] = l > 0 ? TM_EPI_DACT : TM_EPI_NONE;
  }
  rc = 1;
  // (measured on two GPUs, 3,300 rows per rank: beside the peer exchange and the pair kernel the fused backward launch is 3.5 %
  // slower than three launches, while single-GPU small problems gain 3-8 %: slab ranks keep the layer-by-layer backward pass)
  if (tc && !s.slab_api) {
    GemmGroup flat[TM_MAX_HIDDEN * 2 * TM_MAX_ELE];
    for (int l = 0; l < nh; l++)
      for (int i = 0; i < ng; i++) flat[l * ng + i] = all[l][i];
    rc = tm_gemm_tc_launch_multi(c, flat, nh, ng, rowmeta, max_tiles, expect_rows(s), epis, true);
    if (rc < 0) return rc;
    if (rc == 0) tm_trace(c, "backward GEMMs (all layers, one launch)");
  }
  if (rc == 1)
    for (int l = 0; l < nh; l++) {
      if ((rc = tm_launch_gemm(c, all[l], ng, rowmeta, max_tiles, expect_rows(s), epis[l]))) return rc;
      tm_trace(c, "backward GEMM layer");
    }

  return TM_OK;
}
