// C-ABI of libtmolb200.so (see include/tmolb200.h): context, weights, evaluation pipelines and the
// MolEmb / Neighbors.py compatible neighbour-table entry points.
#include "tm_internal.h"
#include <cstdlib>
#include <cuda_fp16.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#define FULL 0xffffffffu

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
void tm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* tm_last_error(void) { return g_err; }
int tm_pdl_enabled() {
  static int on = -1;
  if (on < 0) on = getenv("TM_NO_PDL") ? 0 : 1;
  return on;
}
extern "C" int tm_version(void) { return 100; }
extern "C" int tm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int tm_buf(tm_ctx* c, DevBuf& b, size_t bytes) {
  if (bytes < 256) bytes = 256;
  if (b.cap >= bytes) return TM_OK;
  if (b.p) {
    TM_CUDA(cudaStreamSynchronize(c->stream));
    TM_CUDA(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
  }
  size_t want = bytes + bytes / 8;
  c->alloc_gen++;
  TM_CUDA(cudaMalloc(&b.p, want));
  TM_CUDA(cudaMemsetAsync(b.p, 0, want, c->stream));
  b.cap = want;
  return TM_OK;
}

int tm_host_stage(tm_ctx* c, size_t bytes) {
  if (c->h_cap >= bytes) return TM_OK;
  if (c->h_stage) { TM_CUDA(cudaStreamSynchronize(c->stream)); TM_CUDA(cudaFreeHost(c->h_stage)); c->h_stage = nullptr; c->h_cap = 0; }
  size_t want = bytes + bytes / 4 + 4096;
  c->alloc_gen++;
  TM_CUDA(cudaMallocHost(&c->h_stage, want));
  c->h_cap = want;
  return TM_OK;
}

// ------------------------------------------------------------------------------------------------ params
static int round_up(int v, int m) { return ((v + m - 1) / m) * m; }

static int build_dev_params(tm_ctx* c) {
  const tm_params& p = c->params;
  const tm_model_desc& d = c->desc;
  DevParams& P = c->hp;
  memset(&P, 0, sizeof(P));
  if (d.n_ele < 1 || d.n_ele > TM_MAX_ELE) { tm_set_error("n_ele must be 1..%d", TM_MAX_ELE); return TM_EINVAL; }
  if (p.num_r_Rs < 1 || p.num_r_Rs > 64 || p.num_a_Rs < 1 || p.num_a_Rs > TM_MAX_SYM || p.num_a_As < 1 || p.num_a_As > TM_MAX_SYM) {
    tm_set_error("symmetry-function counts out of range (num_r_Rs<=64, num_a_Rs,num_a_As<=%d)", TM_MAX_SYM);
    return TM_EINVAL;
  }
  if (p.ee_cutoff_on != 0.0) { tm_set_error("EECutoffOn should equal to zero in DSF_elu"); return TM_EINVAL; }   // TFMolInstanceDirect.py:4366
  P.n_ele = d.n_ele;
  P.n_elep = d.n_ele * (d.n_ele + 1) / 2;
  for (int i = 0; i < d.n_ele; i++) {
    P.eles[i] = d.eles[i];
    if (i && d.eles[i] <= d.eles[i - 1]) { tm_set_error("eles must be strictly ascending"); return TM_EINVAL; }
  }
  int l = 0;   // pair channels: upper triangular row-major (TFMolInstanceDirect.py:1262-1267)
  for (int i = 0; i < d.n_ele; i++)
    for (int j = i; j < d.n_ele; j++) {
      P.pair_index[i][j] = (int8_t)l;
      P.pair_index[j][i] = (int8_t)l;
      l++;
    }
  P.nRs_r = p.num_r_Rs; P.nRs_a = p.num_a_Rs; P.nAs = p.num_a_As; P.nsym = p.num_a_As * p.num_a_Rs;
  P.D = d.n_ele * P.nRs_r + P.n_elep * P.nsym;
  P.Dp = round_up(P.D, 128);
  P.r_Rc = (float)p.r_Rc; P.a_Rc = (float)p.a_Rc; P.eta = (float)p.eta; P.zeta = (float)p.zeta;
  P.pi_over_rRc = (float)(3.14159265359 / p.r_Rc);
  P.pi_over_aRc = (float)(3.14159265359 / p.a_Rc);
  P.zeta_pref = (float)pow(2.0, 1.0 - p.zeta);
  P.zeta_is8 = (p.zeta == 8.0);
  for (int s = 0; s < P.nRs_r; s++) P.Rs_r[s] = (float)(p.r_Rc * s / P.nRs_r);       // SetANI1Param
  for (int s = 0; s < P.nRs_a; s++) P.Rs_a[s] = (float)(p.a_Rc * s / P.nRs_a);
  for (int a = 0; a < P.nAs; a++) {
    double th = 2.0 * M_PI * a / P.nAs;
    P.cosA[a] = (float)cos(th);
    P.sinA[a] = (float)sin(th);
  }
  const double B = TM_BOHRPERA;
  double alpha = p.dsf_alpha / B, Rl = p.ee_cutoff_off * B;
  P.R_lr = (float)Rl; P.R_sr = (float)(p.elu_width * B); P.alpha_b = (float)alpha;
  double ZZ = erfc(alpha * Rl) / Rl;
  double YY = 1.1283791671 * alpha * exp(-alpha * Rl * alpha * Rl) / Rl;
  P.Zc = (float)ZZ; P.ZoverR_plus_Y = (float)(ZZ / Rl + YY);
  P.elu_a = (float)p.elu_alpha; P.elu_shift = (float)p.elu_shift;
  P.poly_width_b = (float)(p.poly_width * B);
  P.inv_poly_width_b = (float)(1.0 / (p.poly_width * B));
  {
    // Chebyshev fit (degree 11) of erfcx(x) = exp(x^2) erfc(x) on the LR branch's range, converted to monomials in u
    const int n = 12;
    double lo = std::max(0.0, alpha * p.elu_width * B - 0.05), hi = alpha * Rl + 0.05;
    double mid = 0.5 * (lo + hi), half = 0.5 * (hi - lo);
    double a[n], fx[n];
    for (int k = 0; k < n; k++) {
      double u = cos(M_PI * (k + 0.5) / n), x = mid + half * u;
      fx[k] = exp(x * x) * erfc(x);
    }
    for (int j = 0; j < n; j++) {
      double sacc = 0;
      for (int k = 0; k < n; k++) sacc += fx[k] * cos(M_PI * j * (k + 0.5) / n);
      a[j] = 2.0 * sacc / n;
    }
    a[0] *= 0.5;
    double T0[n] = {0}, T1[n] = {0}, mono[n] = {0}, Tn[n];
    T0[0] = 1.0; T1[1] = 1.0;
    for (int i = 0; i < n; i++) mono[i] += a[0] * T0[i] + a[1] * T1[i];
    for (int j = 2; j < n; j++) {
      for (int i = 0; i < n; i++) Tn[i] = (i > 0 ? 2.0 * T1[i - 1] : 0.0) - T0[i];
      for (int i = 0; i < n; i++) { mono[i] += a[j] * Tn[i]; T0[i] = T1[i]; T1[i] = Tn[i]; }
    }
    double err = 0;
    for (int k = 0; k <= 200; k++) {
      double x = lo + (hi - lo) * k / 200.0, u = (x - mid) / half, pz = mono[n - 1];
      for (int i = n - 2; i >= 0; i--) pz = pz * u + mono[i];
      double ref = exp(x * x) * erfc(x);
      err = std::max(err, fabs(pz - ref) / ref);
    }
    for (int i = 0; i < n; i++) P.erfc_c[i] = (float)mono[i];
    P.erfc_mid = (float)mid; P.erfc_ihalf = (float)(1.0 / half); P.erfc_fit_err = (float)err;
    if (err > 1e-7) { tm_set_error("erfc fit error %.3e too large for this DSFAlpha / cutoff range", err); return TM_EINVAL; }
  }
  for (int i = 0; i < d.n_ele; i++) { P.sqrtC6[i] = (float)sqrt(p.C6[i]); P.Rvdw[i] = (float)p.Rvdw[i]; }
  {
    // pair-kernel constants in Angstrom (see DevParams); every product is formed in double and rounded once
    const double LOG2E = 1.4426950408889634;
    double Rs = p.elu_width * B, ZY = (double)ZZ / Rl + YY;
    double lo = std::max(0.0, alpha * p.elu_width * B - 0.05), hi = alpha * Rl + 0.05;
    double mid = 0.5 * (lo + hi), half = 0.5 * (hi - lo);
    P.pk_rsr2 = (float)(p.elu_width * p.elu_width);
    P.pk_rlr2 = (float)(p.ee_cutoff_off * p.ee_cutoff_off);
    P.pk_cex = (float)(-(alpha * B) * (alpha * B) * LOG2E);
    P.pk_ua = (float)(alpha * B / half);
    P.pk_ub = (float)(-mid / half);
    for (int k = 0; k < 12; k++) P.pk_pc[k] = (float)((double)P.erfc_c[k] / B);
    P.pk_ka = (float)(B * ZY);
    P.pk_kb = (float)(-ZZ - Rl * ZY);
    P.pk_c2 = (float)(1.1283791671 * alpha);
    P.pk_BZY = (float)(B * ZY);
    P.pk_ea = (float)(B * LOG2E);
    P.pk_eb = (float)(-Rs * LOG2E);
    P.pk_ec = (float)(p.elu_shift - p.elu_alpha);
    P.pk_belu = (float)(B * p.elu_alpha);
    P.pk_ta = (float)(B * B / (p.poly_width * B));
    for (int i = 0; i < TM_MAX_ELE; i++)
      for (int j = 0; j < TM_MAX_ELE; j++) {
        double c6 = (i < d.n_ele && j < d.n_ele) ? sqrt(p.C6[i]) * sqrt(p.C6[j]) : 0.0;
        double rs = (i < d.n_ele && j < d.n_ele) ? p.Rvdw[i] + p.Rvdw[j] : 0.0;
        P.pk_c6[i][j] = (float)(c6 / pow(B, 12.0));
        P.pk_rs12[i][j] = (float)(6.0 * pow(rs, 12.0) / pow(B, 24.0));
      }
  }
  P.add_ecc = p.add_ecc; P.activation = p.activation; P.act_alpha = (float)p.sigmoid_alpha;
  P.rr_exact = p.r_Rc; P.ra_exact = p.a_Rc;
  P.skin_on = c->skin > 0.0 ? 1 : 0;
  P.skin = (float)c->skin;
  return TM_OK;
}

extern "C" int tm_set_params(tm_ctx* c, const tm_params* params) {
  if (!c || !params) { tm_set_error("null argument"); return TM_EINVAL; }
  tm_params old = c->params;
  c->cfg_gen++;
  c->params_gen++;
  c->params = *params;
  int rc = build_dev_params(c);
  if (rc) { c->params = old; build_dev_params(c); return rc; }
  return TM_OK;
}

extern "C" tm_ctx* tm_create(int device, const tm_model_desc* desc, const tm_params* params) {
  if (!desc || !params) { tm_set_error("null argument"); return nullptr; }
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    tm_set_error("no CUDA device: libtmolb200 has no CPU fallback");
    return nullptr;
  }
  if (device < 0 || device >= n) { tm_set_error("device %d out of range (have %d)", device, n); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { tm_set_error("cudaSetDevice failed"); return nullptr; }
  tm_ctx* c = new tm_ctx();
  c->device = device;
  c->desc = *desc;
  c->params = *params;
  if (desc->n_hidden < 1 || desc->n_hidden > TM_MAX_HIDDEN) { tm_set_error("n_hidden must be 1..%d", TM_MAX_HIDDEN); delete c; return nullptr; }
  if (build_dev_params(c)) { delete c; return nullptr; }
  c->Hmax = 0;
  for (int l = 0; l < desc->n_hidden; l++) {
    c->Hp[l] = round_up(desc->hidden[l], 128);
    c->Hmax = std::max(c->Hmax, c->Hp[l]);
  }
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { tm_set_error("stream create failed"); delete c; return nullptr; }
  c->stream = c->own_stream;
  cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
  for (int i = 0; i < 12; i++) cudaEventCreate(&c->ev[i]);
  c->graphs_on = getenv("TM_NO_GRAPH") ? 0 : 1;
  c->ev_ok = true;
  memset(&c->last, 0, sizeof(c->last));
  return c;
}

static void free_buf(DevBuf& b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; }

extern "C" void tm_destroy(tm_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->lg.exec) cudaGraphExecDestroy(c->lg.exec);
  tm_gemm_tc_release(c);
  DevBuf* all[] = {&c->b_pos, &c->b_Z, &c->b_cellid, &c->b_rank, &c->b_count, &c->b_cstart, &c->b_sorted, &c->b_satom, &c->b_scan_tmp,
                   &c->b_rowslot, &c->b_rowsidx, &c->b_rowofslot, &c->b_blkcnt, &c->b_rowmeta, &c->b_nbcnt, &c->b_nboff, &c->b_nbr, &c->b_G, &c->b_Gs, &c->b_ypart,
                   &c->b_delta0, &c->b_delta1, &c->b_dG[0], &c->b_dG[1], &c->b_y[0], &c->b_y[1], &c->b_q, &c->b_qs, &c->b_dedq, &c->b_F,
                   &c->b_acc, &c->b_bbox, &c->b_grid, &c->b_flags, &c->b_out, &c->b_molacc, &c->b_natom, &c->b_cntall, &c->b_offall, &c->b_pe, &c->b_pairtab, &c->b_lscan, &c->b_pos0, &c->b_gemm_ready[0], &c->b_gemm_ready[1], &c->b_p2pdone};
  for (DevBuf* b : all) free_buf(*b);
  for (int n = 0; n < 2; n++)
    for (int l = 0; l < TM_MAX_HIDDEN; l++) free_buf(c->b_act[n][l]);
  for (int n = 0; n < 2; n++)
    for (int e = 0; e < TM_MAX_ELE; e++) {
      Net& N = c->nets[n][e];
      for (int l = 0; l < TM_MAX_HIDDEN; l++) {
        if (N.layers[l].W) cudaFree(N.layers[l].W);
        if (N.layers[l].WT) cudaFree(N.layers[l].WT);
        if (N.layers[l].b) cudaFree(N.layers[l].b);
        if (N.layers[l].Ws) cudaFree(N.layers[l].Ws);
        if (N.layers[l].WTs) cudaFree(N.layers[l].WTs);
      }
      if (N.w_out) cudaFree(N.w_out);
    }
  if (c->h_stage) cudaFreeHost(c->h_stage);
  if (c->ev_ok) for (int i = 0; i < 12; i++) cudaEventDestroy(c->ev[i]);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->aux) cudaStreamDestroy(c->aux);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  delete c;
}

extern "C" int tm_set_stream(tm_ctx* c, void* s) {
  if (!c) return TM_EINVAL;
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  c->cfg_gen++;
  return TM_OK;
}
extern "C" int tm_set_gemm_mode(tm_ctx* c, int mode) {
  if (!c || mode < 0 || mode > 4) { tm_set_error("bad gemm mode %d (0 = fp32, 1 = tcgen05 split fp16, 2 = same on CTA pairs, 3 / 4 = mode 1 with 64- / 128-column tiles forced)", mode); return TM_EINVAL; }
  c->gemm_mode = mode;
  c->cfg_gen++;
  return TM_OK;
}
extern "C" int tm_get_gemm_mode(tm_ctx* c) { return c ? c->gemm_mode : TM_EINVAL; }
extern "C" int tm_set_skin(tm_ctx* c, double skin) {
  if (!c || !(skin >= 0.0) || skin > 4.0) { tm_set_error("tm_set_skin: skin must be in [0, 4] Angstrom"); return TM_EINVAL; }
  c->skin = skin;
  c->nl_ok = false;
  c->cfg_gen++;
  return build_dev_params(c);
}
extern "C" int tm_descriptor_width(tm_ctx* c) { return c ? c->hp.D : TM_EINVAL; }
static int check_flags(tm_ctx* c);
extern "C" int tm_sync(tm_ctx* c) {
  if (!c) return TM_EINVAL;
  TM_CUDA(cudaSetDevice(c->device));
  TM_CUDA(cudaStreamSynchronize(c->stream));
  if (c->b_flags.p) return check_flags(c);   // the device-pointer entry points do not synchronise: their capacity / input flags surface here
  return TM_OK;
}

// ------------------------------------------------------------------------------------------------ weights
extern "C" int tm_set_weights(tm_ctx* c, int net, int ele_index, int layer, const double* W, const double* b, int rows, int cols) {
  if (!c || !W || !b) { tm_set_error("null argument"); return TM_EINVAL; }
  if (net < 0 || net > 1 || ele_index < 0 || ele_index >= c->desc.n_ele || layer < 0 || layer > c->desc.n_hidden) {
    tm_set_error("tm_set_weights: index out of range");
    return TM_EINVAL;
  }
  TM_CUDA(cudaSetDevice(c->device));
  c->cfg_gen++;
  int nh = c->desc.n_hidden;
  Net& N = c->nets[net][ele_index];
  if (layer == nh) {
    if (rows != c->desc.hidden[nh - 1] || cols != 1) { tm_set_error("output layer must be [%d][1]", c->desc.hidden[nh - 1]); return TM_EINVAL; }
    int Hp = c->Hp[nh - 1];
    std::vector<float> w((size_t)Hp, 0.f);
    for (int i = 0; i < rows; i++) w[i] = (float)W[i];
    if (!N.w_out) TM_CUDA(cudaMalloc(&N.w_out, (size_t)Hp * 4));
    TM_CUDA(cudaMemcpy(N.w_out, w.data(), (size_t)Hp * 4, cudaMemcpyHostToDevice));
    N.b_out = (float)b[0];
    N.set[nh] = true;
    return TM_OK;
  }
  int K = (layer == 0) ? c->hp.D : c->desc.hidden[layer - 1];
  int Kp = (layer == 0) ? c->hp.Dp : c->Hp[layer - 1];
  int Nn = c->desc.hidden[layer], Np = c->Hp[layer];
  if (rows != K || cols != Nn) { tm_set_error("layer %d must be [%d][%d], got [%d][%d]", layer, K, Nn, rows, cols); return TM_EINVAL; }
  Layer& L = N.layers[layer];
  L.K = K; L.N = Nn; L.Kp = Kp; L.Np = Np;
  std::vector<float> w((size_t)Kp * Np, 0.f), wt((size_t)Np * Kp, 0.f), bb((size_t)Np, 0.f);
  for (int k = 0; k < K; k++)
    for (int n = 0; n < Nn; n++) {
      float v = (float)W[(size_t)k * Nn + n];
      w[(size_t)k * Np + n] = v;
      wt[(size_t)n * Kp + k] = v;
    }
  for (int n = 0; n < Nn; n++) bb[n] = (float)b[n];
  if (!L.W) TM_CUDA(cudaMalloc(&L.W, w.size() * 4));
  if (!L.WT) TM_CUDA(cudaMalloc(&L.WT, wt.size() * 4));
  if (!L.b) TM_CUDA(cudaMalloc(&L.b, bb.size() * 4));
  TM_CUDA(cudaMemcpy(L.W, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  TM_CUDA(cudaMemcpy(L.WT, wt.data(), wt.size() * 4, cudaMemcpyHostToDevice));
  TM_CUDA(cudaMemcpy(L.b, bb.data(), bb.size() * 4, cudaMemcpyHostToDevice));
  // split-fp16 planes for the tcgen05 path: hi = rn_f16(x), lo = rn_f16((x - hi) * 2048)   (tm_gemm_tc.cu)
  bool range_ok = true;
  auto split = [&range_ok](const std::vector<float>& src, std::vector<__half>& dst) {
    size_t n = src.size();
    dst.resize(2 * n);
    for (size_t i = 0; i < n; i++) {
      if (!(fabsf(src[i]) <= 65000.f)) range_ok = false;
      __half hi = __float2half_rn(src[i]);
      dst[i] = hi;
      dst[n + i] = __float2half_rn((src[i] - __half2float(hi)) * 2048.0f);
    }
  };
  std::vector<__half> ws, wts;
  split(w, ws);
  split(wt, wts);
  if (!range_ok) { tm_set_error("tm_set_weights: a weight is outside the fp16 range of the tensor-core path"); return TM_EINVAL; }
  if (!L.Ws) TM_CUDA(cudaMalloc(&L.Ws, ws.size() * 2));
  if (!L.WTs) TM_CUDA(cudaMalloc(&L.WTs, wts.size() * 2));
  TM_CUDA(cudaMemcpy(L.Ws, ws.data(), ws.size() * 2, cudaMemcpyHostToDevice));
  TM_CUDA(cudaMemcpy(L.WTs, wts.data(), wts.size() * 2, cudaMemcpyHostToDevice));
  N.set[layer] = true;
  return TM_OK;
}

static int check_weights(tm_ctx* c) {
  for (int n = 0; n < 2; n++)
    for (int e = 0; e < c->desc.n_ele; e++)
      for (int l = 0; l <= c->desc.n_hidden; l++)
        if (!c->nets[n][e].set[l]) {
          tm_set_error("weights not set: net %d element index %d layer %d (call tm_set_weights)", n, e, l);
          return TM_ESTATE;
        }
  return TM_OK;
}

// ------------------------------------------------------------------------------------------------ small kernels
// per-molecule Ebp (fp32 GEMM mode only: the tensor-core forward pass sums it in k_y_reduce)
__global__ void k_ebp(const float* __restrict__ y, const int32_t* __restrict__ rowslot, int64_t nrows, int64_t maxnatom,
                      double* __restrict__ molacc) {
  TM_PDL_PROLOGUE;
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int slot = (r < nrows) ? rowslot[r] : -1;
  double v = 0.0;
  int m = -1;
  if (slot >= 0) {
    v = (double)y[r];
    m = (int)(slot / maxnatom);
  }
  // warp-aggregate when the whole warp belongs to one molecule
  int m0 = __shfl_sync(FULL, m, 0);
  bool uniform = __all_sync(FULL, m == m0 || m < 0);
  if (uniform) {
    int mm = m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v += __shfl_xor_sync(FULL, v, o);
      mm = max(mm, __shfl_xor_sync(FULL, mm, o));
    }
    if ((threadIdx.x & 31) == 0 && mm >= 0) atomicAdd(&molacc[16 * mm + 1], v);
  } else if (m >= 0) {
    atomicAdd(&molacc[16 * m + 1], v);
  }
}

// pack the outputs as doubles: [Etot nmol][Ebp][Ecc][Evdw][dipole 3nmol][Ebp_atom nq][charge nq][grad 3nq]
// (one launch: the molecule sums are complete, Ebp_atom comes from the energy net's rows through rowofslot)
__global__ void k_pack_all(const double* __restrict__ molacc, int64_t nmol, int add_ecc, const float* __restrict__ y_e,
                           const int32_t* __restrict__ rowofslot, const double* __restrict__ q_slot, const float* __restrict__ F, int64_t nq,
                           int do_force, double* __restrict__ out) {
  TM_PDL_PROLOGUE;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (int64_t m = t0; m < nmol; m += stride) {
    double ebp = molacc[16 * m + 1], ecc = add_ecc ? molacc[16 * m + 2] : 0.0, evdw = molacc[16 * m + 3];
    out[m] = ebp + ecc + evdw;                 // TFMolInstanceDirect.py:5213-5215
    out[nmol + m] = ebp;
    out[2 * nmol + m] = ecc;
    out[3 * nmol + m] = evdw;
    out[4 * nmol + 3 * m] = molacc[16 * m + 6];
    out[4 * nmol + 3 * m + 1] = molacc[16 * m + 7];
    out[4 * nmol + 3 * m + 2] = molacc[16 * m + 8];
  }
  double* ebp_atom = out + 7 * nmol;
  double* charge = ebp_atom + nq;
  double* grad = charge + nq;
  for (int64_t t = t0; t < nq; t += stride) {
    int row = rowofslot[t];
    ebp_atom[t] = (row >= 0) ? (double)y_e[row] : 0.0;
    charge[t] = q_slot[t];
  }
  if (do_force)
    for (int64_t t = t0; t < 3 * nq; t += stride) grad[t] = (double)F[t];
}
__global__ void k_f2d(const float* __restrict__ in, double* __restrict__ out, int64_t n) {
  TM_PDL_PROLOGUE;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) out[t] = (double)in[t];
}
__global__ void k_desc_out(const float* __restrict__ G, const int32_t* __restrict__ rowofslot, int64_t nq, int D, int Dp, float* __restrict__ out) {
  TM_PDL_PROLOGUE;
  int64_t slot = blockIdx.x;
  if (slot >= nq) return;
  int row = rowofslot[slot];
  for (int d = threadIdx.x; d < D; d += blockDim.x) out[slot * D + d] = (row >= 0) ? G[(int64_t)row * Dp + d] : 0.f;
}
__global__ void k_set_i32(int32_t* p, int64_t n, int32_t v) {
  TM_PDL_PROLOGUE;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) p[t] = v;
}

static inline int nblk(int64_t n, int per = 256, int cap = 148 * 8) {
  int64_t b = (n + per - 1) / per;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

// ------------------------------------------------------------------------------------------------ pipeline
struct OutLayout {
  int64_t nmol, nq, off_ebp_atom, off_charge, off_grad, total;
};
static OutLayout out_layout(int64_t nmol, int64_t nq) {
  OutLayout o;
  o.nmol = nmol; o.nq = nq;
  o.off_ebp_atom = 7 * nmol;
  o.off_charge = o.off_ebp_atom + nq;
  o.off_grad = o.off_charge + nq;
  o.total = o.off_grad + 3 * nq;
  return o;
}

static SysView make_view(tm_ctx* c, int64_t nslots, int64_t nmol, int64_t maxnatom, int64_t nreal, int periodic, int64_t ncent_max) {
  SysView s;
  s.nslots = nslots; s.nmol = nmol; s.maxnatom = maxnatom; s.nreal = nreal; s.periodic = periodic;
  s.ncent_max = ncent_max;
  s.nrows = ncent_max + (int64_t)TM_ROW_TILE * c->hp.n_ele;
  s.ncells_cap = nslots + 1024;
  s.slab_rank = 0; s.slab_world = 1; s.slab_api = 0;
  s.slab_g[0] = s.slab_g[1] = s.slab_g[2] = 0.0;
  s.window_on = 0; s.win_lo = 0.0; s.win_hi = 0.0;
  s.win_ntess = 0; s.win_ilo = -1000; s.win_ihi = 1000;
  s.grid_host = 0;
  memset(&s.hgrid, 0, sizeof(s.hgrid));
  s.lat_bin = 0; s.lat_ntess = 0; s.xyz_real = nullptr; s.Z_real = nullptr;
  memset(&s.lat, 0, sizeof(s.lat));
  for (int d = 0; d < 9; d++) s.ginv[d] = 0.0;
  for (int d = 0; d < 3; d++) { s.wlo[d] = 0.0; s.whi[d] = 0.0; }
  return s;
}

// TM_TRACE=1 (debugging): synchronise after every stage and report it on stderr, so that a stage that never finishes or
// faults is named.  Off by default (one getenv per process); never active during graph capture.
void tm_trace(tm_ctx* c, const char* what) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("TM_TRACE"); on = (e && atoi(e) != 0) ? 1 : 0; }
  if (!on) return;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(c->stream, &st);
  if (st != cudaStreamCaptureStatusNone) return;
  fprintf(stderr, "[tm_trace] %s ...", what);
  fflush(stderr);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  fprintf(stderr, " %s\n", e == cudaSuccess ? "ok" : cudaGetErrorString(e));
  fflush(stderr);
}

// stages up to and including the nets' forward pass (pos/Z/inv_n already on the device)
static int stage_a(tm_ctx* c, const SysView& s, bool reuse = false) {
  int rc;
  int64_t nq = s.periodic ? s.nreal : s.nslots;
  if ((rc = tm_buf(c, c->b_flags, 64))) return rc;
  if ((rc = tm_buf(c, c->b_molacc, (size_t)s.nmol * 16 * 8))) return rc;
  if ((rc = tm_buf(c, c->b_F, (size_t)nq * 3 * 4))) return rc;
  TM_CUDA(cudaMemsetAsync(c->b_flags.p, 0, 32, c->stream));   // words 8.. are sticky (read and cleared by check_flags)
  TM_CUDA(cudaMemsetAsync(c->b_molacc.p, 0, (size_t)s.nmol * 16 * 8, c->stream));
  TM_CUDA(cudaMemsetAsync(c->b_F.p, 0, (size_t)nq * 3 * 4, c->stream));
  cudaEventRecord(c->ev[1], c->stream);
  tm_trace(c, "inputs");
  if (reuse) {
    if ((rc = tm_launch_lattice_refresh(c, s))) return rc;
    tm_trace(c, "position refresh (neighbour rows reused)");
  } else if (s.lat_bin) {
    if ((rc = tm_launch_lattice_bin(c, s))) return rc;
    tm_trace(c, "windowed lattice binning + rows");
  } else {
    if ((rc = tm_launch_nlist_build(c, s, c->params.r_Rc + c->skin))) return rc;
    tm_trace(c, "cell list");
    if ((rc = tm_launch_rows(c, s))) return rc;
    tm_trace(c, "rows");
  }
  if (!reuse) {
    if ((rc = tm_launch_neighbours(c, s))) return rc;
    tm_trace(c, "neighbour rows");
  }
  cudaEventRecord(c->ev[2], c->stream);
  if ((rc = tm_launch_desc(c, s))) return rc;
  tm_trace(c, "descriptors");
  cudaEventRecord(c->ev[3], c->stream);
  if ((rc = tm_launch_mlp_forward(c, s))) return rc;
  tm_trace(c, "nets forward");
  cudaEventRecord(c->ev[4], c->stream);
  return TM_OK;
}

static int stage_b(tm_ctx* c, const SysView& s, int flags) {
  int rc;
  if ((rc = tm_launch_charges(c, s))) return rc;
  tm_trace(c, "charges");
  if ((rc = tm_launch_pair(c, s, flags))) return rc;
  tm_trace(c, "pair kernel");
  cudaEventRecord(c->ev[5], c->stream);
  return TM_OK;
}

// Small problems (slab ranks, small cells): the backward GEMMs leave most SMs idle in their second tile round and the
// pair kernel has long per-warp chains; neither needs the other's result, so the backward pass goes to the side stream
// right after the forward pass and the two fill each other's gaps.  Large problems keep the single stream (both kernels
// fill the machine on their own, and the stage timings stay separable).  TM_NO_OVERLAP=1 disables it.
static bool overlap_backward(const tm_ctx* c, const SysView& s) {
  static int off = -1;
  if (off < 0) off = getenv("TM_NO_OVERLAP") ? 1 : 0;
  if (off || !c->aux) return false;
  return tm_expected_centres(s) <= 8192;
}
static int fork_backward(tm_ctx* c, const SysView& s) {
  TM_CUDA(cudaEventRecord(c->ev_fork, c->stream));
  TM_CUDA(cudaStreamWaitEvent(c->aux, c->ev_fork, 0));
  cudaStream_t keep = c->stream;
  c->stream = c->aux;
  int rc = tm_launch_mlp_backward(c, s);
  c->stream = keep;
  if (rc) return rc;
  TM_CUDA(cudaEventRecord(c->ev_join, c->aux));
  c->fork_open = true;
  return TM_OK;
}
static int join_backward(tm_ctx* c) {
  if (c->fork_open) {
    TM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    c->fork_open = false;
  }
  return TM_OK;
}

static int stage_c(tm_ctx* c, const SysView& s, int flags) {
  int rc;
  if (flags & TM_F_FORCE) {
    if (c->fork_open) {
      if ((rc = join_backward(c))) return rc;
    } else if ((rc = tm_launch_mlp_backward(c, s))) {
      return rc;
    }
    tm_trace(c, "nets backward");
    cudaEventRecord(c->ev[6], c->stream);
    if ((rc = tm_launch_force(c, s, flags))) return rc;
    tm_trace(c, "force kernel");
  } else {
    cudaEventRecord(c->ev[6], c->stream);
  }
  cudaEventRecord(c->ev[7], c->stream);
  return TM_OK;
}

// energies / per-atom outputs into the packed double buffer
static int stage_pack(tm_ctx* c, const SysView& s, int flags, const OutLayout& o) {
  int rc;
  if ((rc = tm_buf(c, c->b_out, (size_t)o.total * 8))) return rc;
  double* out = (double*)c->b_out.p;
  if (!c->y_fused) {
    TM_LAUNCH(k_ebp, (int)((s.nrows + 255) / 256), 256, 0, c->stream, (const float*)c->b_y[TM_NET_ENERGY].p, (const int32_t*)c->b_rowslot.p, s.nrows, s.maxnatom,
                                                              (double*)c->b_molacc.p);
    c->launches++;
  }
  const int do_force = (flags & TM_F_FORCE) ? 1 : 0;
  TM_LAUNCH(k_pack_all, nblk(std::max<int64_t>(s.nmol, (do_force ? 3 : 1) * o.nq)), 256, 0, c->stream, 
      (const double*)c->b_molacc.p, s.nmol, c->hp.add_ecc, (const float*)c->b_y[TM_NET_ENERGY].p, (const int32_t*)c->b_rowofslot.p,
      (const double*)c->b_q.p + o.nq, (const float*)c->b_F.p, o.nq, do_force, out);
  c->launches++;
  TM_CUDA(cudaGetLastError());
  return TM_OK;
}

static int finish_timings(tm_ctx* c, const SysView& s) {
  tm_timings& t = c->last;
  memset(&t, 0, sizeof(t));
  auto el = [&](int a, int b) { float ms = 0.f; cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]); return ms; };
  t.h2d = el(0, 1); t.nlist = el(1, 2); t.desc = el(2, 3); t.mlp_fwd = el(3, 4); t.pair = el(4, 5); t.mlp_bwd = el(5, 6);
  t.force = el(6, 7); t.d2h = el(7, 8); t.total = el(0, 8);
  t.n_slots = s.nslots;
  t.launches = c->launches;
  return TM_OK;
}

static int check_flags(tm_ctx* c) {
  int32_t f[2] = {0, 0};
  int32_t sticky[2] = {0, 0};
  TM_CUDA(cudaMemcpyAsync(f, c->b_flags.p, 8, cudaMemcpyDeviceToHost, c->stream));
  TM_CUDA(cudaMemcpyAsync(sticky, (char*)c->b_flags.p + 32, 8, cudaMemcpyDeviceToHost, c->stream));
  TM_CUDA(cudaStreamSynchronize(c->stream));
  if (sticky[0]) {      // raised by a step that may have been replayed from a graph many steps ago (TM_F_REUSE_NLIST)
    TM_CUDA(cudaMemsetAsync((char*)c->b_flags.p + 32, 0, 8, c->stream));
    f[0] |= sticky[0];
  }
  c->last_flags = f[0];
  if (f[0] & 2) { tm_set_error("more than %d radial neighbours of one centre", TM_NB_STRIDE); return TM_ECAP; }
  if (f[0] & 32) { tm_set_error("slab exchange: a peer rank never signalled (timed out after ~8 s)"); return TM_ECUDA; }
  if (f[0] & 8) { tm_set_error("coordinates are not wrapped into the cell (apply Lattice.ModuloLattice before tm_eval_lattice)"); return TM_EINVAL; }
  if (f[0] & 4) { tm_set_error("more than %d neighbours inside the angular cutoff of one centre", TM_ANG_CAP); return TM_ECAP; }
  if (f[0] & 256) { tm_set_error("fused GEMM: a layer dependency was never published (timed out)"); return TM_ECUDA; }
  if (f[0] & 128) { tm_set_error("TM_F_REUSE_NLIST: an atom moved more than skin / 2 since the neighbour rows were built (rebuild more often or raise the skin)"); return TM_ESTATE; }
  if (f[0] & 64) { tm_set_error("a slab rank owns more centres than its row allocation (density far from uniform)"); return TM_ECAP; }
  return TM_OK;
}

static int validate_Z(tm_ctx* c, const int32_t* Z, int64_t n) {
  bool known[256] = {};
  for (int k = 0; k < c->desc.n_ele; k++)
    if (c->desc.eles[k] > 0 && c->desc.eles[k] < 256) known[c->desc.eles[k]] = true;
  for (int64_t i = 0; i < n; i++) {
    int z = Z[i];
    if (z <= 0) continue;
    if (z > 255 || !known[z]) { tm_set_error("atomic number %d (slot %lld) is not in the model's element list", z, (long long)i); return TM_EINVAL; }
  }
  return TM_OK;
}

// which per-atom blocks of the packed output the caller asked for (bit 0 Ebp_atom, 1 charge, 2 gradient): only those
// cross the bus (the reference's periodic callback returns Etotal and the force, TFMolManage.py:1353-1358)
static int out_mask(int flags, const tm_outputs* out) {
  return (out->Ebp_atom ? 1 : 0) | (out->charge ? 2 : 0) | ((out->gradient && (flags & TM_F_FORCE)) ? 4 : 0);
}
// page-locked host memory (cudaMallocHost / cudaHostRegister / a torch pinned tensor): the copy engine reads and writes
// it directly, so a caller who keeps its arrays there skips the staging memcpy on both sides
static bool host_pinned(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}
// D2H of the energies (always) and the requested blocks, each to its own offset of the pinned staging buffer, or
// straight into the caller's array where `direct` names one (page-locked; blocks 0 Ebp_atom, 1 charge, 2 gradient)
static int copy_packed_d2h(tm_ctx* c, const OutLayout& o, int mask, void* const* direct = nullptr) {
  char* hs = (char*)c->h_stage;
  const char* d = (const char*)c->b_out.p;
  TM_CUDA(cudaMemcpyAsync(hs, d, (size_t)7 * o.nmol * 8, cudaMemcpyDeviceToHost, c->stream));
  if (direct && (direct[0] || direct[1] || direct[2])) {
    const int64_t off[3] = {o.off_ebp_atom, o.off_charge, o.off_grad};
    const int64_t cnt[3] = {o.nq, o.nq, 3 * o.nq};
    for (int b = 0; b < 3; b++)
      if (mask & (1 << b))
        TM_CUDA(cudaMemcpyAsync(direct[b] ? (char*)direct[b] : hs + off[b] * 8, d + off[b] * 8, (size_t)cnt[b] * 8, cudaMemcpyDeviceToHost, c->stream));
    return TM_OK;
  }
  if (mask == 7) {   // everything: one copy
    TM_CUDA(cudaMemcpyAsync(hs + o.off_ebp_atom * 8, d + o.off_ebp_atom * 8, (size_t)(o.total - o.off_ebp_atom) * 8, cudaMemcpyDeviceToHost, c->stream));
    return TM_OK;
  }
  if (mask & 1) TM_CUDA(cudaMemcpyAsync(hs + o.off_ebp_atom * 8, d + o.off_ebp_atom * 8, (size_t)o.nq * 8, cudaMemcpyDeviceToHost, c->stream));
  if (mask & 2) TM_CUDA(cudaMemcpyAsync(hs + o.off_charge * 8, d + o.off_charge * 8, (size_t)o.nq * 8, cudaMemcpyDeviceToHost, c->stream));
  if (mask & 4) TM_CUDA(cudaMemcpyAsync(hs + o.off_grad * 8, d + o.off_grad * 8, (size_t)3 * o.nq * 8, cudaMemcpyDeviceToHost, c->stream));
  return TM_OK;
}

static int copy_out(tm_ctx* c, int flags, const OutLayout& o, tm_outputs* out, int64_t charge_tile_to, size_t bytes, size_t dbytes,
                    void* const* direct = nullptr);

// copy the packed device outputs to the user's arrays
static int deliver(tm_ctx* c, const SysView& s, int flags, const OutLayout& o, tm_outputs* out, int64_t charge_tile_to) {
  int rc;
  size_t bytes = (size_t)o.total * 8;
  size_t dbytes = (flags & TM_F_DESCRIPTORS) && out->descriptors ? (size_t)o.nq * c->hp.D * 4 : 0;
  if ((rc = tm_host_stage(c, bytes + dbytes))) return rc;
  if ((rc = copy_packed_d2h(c, o, out_mask(flags, out)))) return rc;
  if (dbytes) {
    if ((rc = tm_buf(c, c->b_acc, dbytes))) return rc;
    TM_LAUNCH(k_desc_out, (unsigned)o.nq, 128, 0, c->stream, (const float*)c->b_G.p, (const int32_t*)c->b_rowofslot.p, o.nq, c->hp.D, c->hp.Dp, (float*)c->b_acc.p);
    c->launches++;
    TM_CUDA(cudaMemcpyAsync((char*)c->h_stage + bytes, c->b_acc.p, dbytes, cudaMemcpyDeviceToHost, c->stream));
  }
  cudaEventRecord(c->ev[8], c->stream);
  if ((rc = check_flags(c))) return rc;   // synchronises
  if ((rc = copy_out(c, flags, o, out, charge_tile_to, bytes, dbytes))) return rc;
  return finish_timings(c, s);
}

// host side of a delivery: the packed outputs are in the pinned staging buffer
static int copy_out(tm_ctx* c, int flags, const OutLayout& o, tm_outputs* out, int64_t charge_tile_to, size_t bytes, size_t dbytes,
                    void* const* direct) {
  const double* h = (const double*)c->h_stage;
  int64_t nm = o.nmol;
  for (int64_t m = 0; m < nm; m++)
    if (!std::isfinite(h[m])) {
      tm_set_error("non-finite energy for molecule %lld (bad input%s)", (long long)m,
                   c->gemm_mode != TM_GEMM_FP32 ? ", or an MLP activation outside the fp16 range of gemm mode 1: use mode 0" : "");
      return TM_ECAP;
    }
  if (out->Etotal) memcpy(out->Etotal, h, nm * 8);
  if (out->Ebp) memcpy(out->Ebp, h + nm, nm * 8);
  if (out->Ecc) memcpy(out->Ecc, h + 2 * nm, nm * 8);
  if (out->Evdw) memcpy(out->Evdw, h + 3 * nm, nm * 8);
  if (out->dipole) memcpy(out->dipole, h + 4 * nm, 3 * nm * 8);
  if (out->Ebp_atom && !(direct && direct[0])) memcpy(out->Ebp_atom, h + o.off_ebp_atom, o.nq * 8);
  if (out->charge && !(direct && direct[1])) {
    int64_t done = 0;
    while (done < charge_tile_to) {   // periodic: tile q over the image blocks (TFMolInstanceDirect.py:5892-5893)
      int64_t n = std::min(o.nq, charge_tile_to - done);
      memcpy(out->charge + done, h + o.off_charge, n * 8);
      done += n;
    }
  }
  if (out->gradient && (flags & TM_F_FORCE) && !(direct && direct[2])) memcpy(out->gradient, h + o.off_grad, 3 * o.nq * 8);
  if (dbytes) memcpy(out->descriptors, (const char*)c->h_stage + bytes, dbytes);
  return TM_OK;
}

static int upload_inv_n(tm_ctx* c, const double* inv_n, int64_t nmol) {
  int rc;
  if ((rc = tm_buf(c, c->b_natom, (size_t)nmol * 8))) return rc;
  TM_CUDA(cudaMemcpyAsync(c->b_natom.p, inv_n, (size_t)nmol * 8, cudaMemcpyHostToDevice, c->stream));
  return TM_OK;
}

static int run_all(tm_ctx* c, const SysView& s, int flags, const OutLayout& o, bool reuse = false) {
  int rc;
  if ((rc = stage_a(c, s, reuse))) return rc;
  if ((flags & TM_F_FORCE) && overlap_backward(c, s) && (rc = fork_backward(c, s))) return rc;
  if ((rc = stage_b(c, s, flags))) return rc;
  if ((rc = stage_c(c, s, flags))) return rc;
  return stage_pack(c, s, flags, o);
}

extern "C" int tm_eval(tm_ctx* c, const double* xyzs, const int32_t* Zs, int64_t nmol, int64_t maxnatom, const int64_t* natom, int flags,
                       tm_outputs* out) {
  if (c) c->nl_ok = false;
  if (!c || !xyzs || !Zs || !natom || !out || nmol < 1 || maxnatom < 1) { tm_set_error("tm_eval: bad argument"); return TM_EINVAL; }
  int rc;
  TM_CUDA(cudaSetDevice(c->device));
  if ((rc = check_weights(c))) return rc;
  int64_t nslots = nmol * maxnatom;
  if ((rc = validate_Z(c, Zs, nslots))) return rc;
  c->launches = 0;
  int64_t ncent = 0;
  // staging: xyz | Z (masked by natom) | inv_n
  size_t bx = (size_t)nslots * 24, bz = (size_t)nslots * 4, bn = (size_t)nmol * 8;
  if ((rc = tm_host_stage(c, bx + bz + bn + 64))) return rc;
  char* hs = (char*)c->h_stage;
  memcpy(hs, xyzs, bx);
  int32_t* hz = (int32_t*)(hs + bx);
  double* hn = (double*)(hs + bx + bz);
  for (int64_t m = 0; m < nmol; m++) {
    if (natom[m] < 0 || natom[m] > maxnatom) { tm_set_error("natom[%lld] out of range", (long long)m); return TM_EINVAL; }
    for (int64_t a = 0; a < maxnatom; a++) {
      int32_t z = (a < natom[m]) ? Zs[m * maxnatom + a] : 0;
      hz[m * maxnatom + a] = z;
      if (z > 0) ncent++;
    }
    hn[m] = natom[m] > 0 ? 1.0 / (double)natom[m] : 0.0;
  }
  if ((rc = tm_buf(c, c->b_pos, bx))) return rc;
  if ((rc = tm_buf(c, c->b_Z, bz))) return rc;
  c->timings_final = false;
  cudaEventRecord(c->ev[0], c->stream);
  TM_CUDA(cudaMemcpyAsync(c->b_pos.p, hs, bx, cudaMemcpyHostToDevice, c->stream));
  TM_CUDA(cudaMemcpyAsync(c->b_Z.p, hz, bz, cudaMemcpyHostToDevice, c->stream));
  if ((rc = upload_inv_n(c, hn, nmol))) return rc;
  SysView s = make_view(c, nslots, nmol, maxnatom, 0, 0, ncent);
  OutLayout o = out_layout(nmol, nslots);
  if ((rc = run_all(c, s, flags, o))) return rc;
  rc = deliver(c, s, flags, o, out, nslots);
  c->last.n_centres = ncent;
  return rc;
}

// 1 / natom per molecule from the zero-padded atomic numbers (device form of upload_inv_n)
__global__ void k_inv_n_from_Z(const int32_t* __restrict__ Z, int64_t nmol, int64_t maxnatom, double* __restrict__ inv_n) {
  TM_PDL_PROLOGUE;
  int64_t m = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (m >= nmol) return;
  int cnt = 0;
  for (int64_t a = lane; a < maxnatom; a += 32) cnt += Z[m * maxnatom + a] > 0 ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) inv_n[m] = cnt > 0 ? 1.0 / (double)cnt : 0.0;
}

// Device-resident form of tm_eval (see include/tmolb200.h): no host copy, no synchronisation, capturable.
extern "C" int tm_eval_dev(tm_ctx* c, const double* xyz_dev, const int32_t* Z_dev, int64_t nmol, int64_t maxnatom, int flags,
                           double* e_dev, double* grad_dev, double* charge_dev) {
  if (c) c->nl_ok = false;
  if (!c || !xyz_dev || !Z_dev || nmol < 1 || maxnatom < 1) { tm_set_error("tm_eval_dev: bad argument"); return TM_EINVAL; }
  int rc;
  TM_CUDA(cudaSetDevice(c->device));
  if ((rc = check_weights(c))) return rc;
  const int64_t nslots = nmol * maxnatom;
  c->launches = 0;
  c->timings_final = false;
  cudaEventRecord(c->ev[0], c->stream);
  if ((rc = tm_buf(c, c->b_pos, (size_t)nslots * 24))) return rc;
  if ((rc = tm_buf(c, c->b_Z, (size_t)nslots * 4))) return rc;
  if ((rc = tm_buf(c, c->b_natom, (size_t)nmol * 8))) return rc;
  TM_CUDA(cudaMemcpyAsync(c->b_pos.p, xyz_dev, (size_t)nslots * 24, cudaMemcpyDeviceToDevice, c->stream));
  TM_CUDA(cudaMemcpyAsync(c->b_Z.p, Z_dev, (size_t)nslots * 4, cudaMemcpyDeviceToDevice, c->stream));
  TM_LAUNCH(k_inv_n_from_Z, (unsigned)((nmol * 32 + 255) / 256), 256, 0, c->stream, (const int32_t*)c->b_Z.p, nmol, maxnatom, (double*)c->b_natom.p);
  c->launches++;
  SysView s = make_view(c, nslots, nmol, maxnatom, 0, 0, nslots);   // every slot may be a centre
  OutLayout o = out_layout(nmol, nslots);
  if ((rc = run_all(c, s, flags, o))) return rc;
  const double* out = (const double*)c->b_out.p;
  if (e_dev) TM_CUDA(cudaMemcpyAsync(e_dev, out, (size_t)4 * nmol * 8, cudaMemcpyDeviceToDevice, c->stream));
  if (grad_dev && (flags & TM_F_FORCE)) TM_CUDA(cudaMemcpyAsync(grad_dev, out + o.off_grad, (size_t)3 * nslots * 8, cudaMemcpyDeviceToDevice, c->stream));
  if (charge_dev) TM_CUDA(cudaMemcpyAsync(charge_dev, out + o.off_charge, (size_t)nslots * 8, cudaMemcpyDeviceToDevice, c->stream));
  cudaEventRecord(c->ev[8], c->stream);
  c->cur_nslots = nslots;
  return TM_OK;
}

extern "C" int tm_eval_images(tm_ctx* c, const double* xyz_tess, const int32_t* Z_tess, int64_t ntess_atoms, int64_t nreal, int flags,
                              tm_outputs* out) {
  if (c) c->nl_ok = false;
  if (!c || !xyz_tess || !Z_tess || !out || nreal < 1 || ntess_atoms < nreal) { tm_set_error("tm_eval_images: bad argument"); return TM_EINVAL; }
  int rc;
  TM_CUDA(cudaSetDevice(c->device));
  if ((rc = check_weights(c))) return rc;
  if ((rc = validate_Z(c, Z_tess, ntess_atoms))) return rc;
  c->launches = 0;
  size_t bx = (size_t)ntess_atoms * 24, bz = (size_t)ntess_atoms * 4;
  if ((rc = tm_host_stage(c, bx + bz + 64))) return rc;
  char* hs = (char*)c->h_stage;
  memcpy(hs, xyz_tess, bx);
  memcpy(hs + bx, Z_tess, bz);
  double inv = 1.0 / (double)nreal;   // natom is fed as nreal (TFMolManage.py:1342)
  memcpy(hs + bx + bz, &inv, 8);
  if ((rc = tm_buf(c, c->b_pos, bx))) return rc;
  if ((rc = tm_buf(c, c->b_Z, bz))) return rc;
  c->timings_final = false;
  cudaEventRecord(c->ev[0], c->stream);
  TM_CUDA(cudaMemcpyAsync(c->b_pos.p, hs, bx, cudaMemcpyHostToDevice, c->stream));
  TM_CUDA(cudaMemcpyAsync(c->b_Z.p, hs + bx, bz, cudaMemcpyHostToDevice, c->stream));
  if ((rc = upload_inv_n(c, (const double*)(hs + bx + bz), 1))) return rc;
  SysView s = make_view(c, ntess_atoms, 1, ntess_atoms, nreal, 1, nreal);
  OutLayout o = out_layout(1, nreal);
  if ((rc = run_all(c, s, flags, o))) return rc;
  rc = deliver(c, s, flags, o, out, ntess_atoms);
  c->last.n_centres = nreal;
  return rc;
}

static int64_t tess_count(int64_t nreal, int ntess) {
  int side = 2 * ntess + 1;
  return (int64_t)side * side * side * nreal;
}

// device-side tessellation; xyz_dev / Z_dev are DEVICE pointers to the primitive cell

static int prepare_lattice_full(tm_ctx* c, const double* xyz_dev, const int32_t* Z_dev, int64_t nreal, const double* lattice, int ntess, SysView* sv,
                           int ilo = -1000, int ihi = 1000) {
  int rc;
  if (ntess < 1 || ntess > 8) { tm_set_error("ntess must be 1..8"); return TM_EINVAL; }
  int64_t nslots = tess_count(nreal, ntess);
  if ((rc = tm_buf(c, c->b_pos, (size_t)nslots * 24))) return rc;
  if ((rc = tm_buf(c, c->b_Z, (size_t)nslots * 4))) return rc;
  if ((rc = tm_buf(c, c->b_natom, 8))) return rc;
  // lattice and 1/natom travel as kernel arguments (no pageable copy: the call sequence stays CUDA-graph capturable);
  // k_tessellate stores 1/natom into b_natom for the charge kernels
  LatArgs la;
  memcpy(la.v, lattice, 72);
  la.v[9] = 1.0 / (double)nreal;
  if ((rc = tm_launch_tessellate(c, xyz_dev, Z_dev, nreal, la, ntess, ilo, ihi))) return rc;
  *sv = make_view(c, nslots, 1, nslots, nreal, 1, nreal);
  // inverse lattice first row (for slab ownership): frac_a = pos . g
  const double* L = lattice;
  double det = L[0] * (L[4] * L[8] - L[5] * L[7]) - L[1] * (L[3] * L[8] - L[5] * L[6]) + L[2] * (L[3] * L[7] - L[4] * L[6]);
  if (fabs(det) < 1e-12) { tm_set_error("singular lattice"); return TM_EINVAL; }
  sv->slab_g[0] = (L[4] * L[8] - L[5] * L[7]) / det;
  sv->slab_g[1] = -(L[3] * L[8] - L[5] * L[6]) / det;   // g = first COLUMN of inv(L)  (x = f L  =>  f = x inv(L)), i.e. cofactors of row 0
  sv->slab_g[2] = (L[3] * L[7] - L[4] * L[6]) / det;
  return TM_OK;
}

// Lattice path: the binned atoms (all images, or the slab window of them) fill a parallelepiped that is known on the
// host, so the cell grid is laid out here instead of by a bounding-box pass over the slots (same rules as
// k_grid_params).  Atoms outside it (unwrapped input) raise flag 8 in k_cell_count.
static void host_grid(tm_ctx* c, SysView* sv, const double* L, int ntess) {
  double flo[3], fhi[3];
  for (int d = 0; d < 3; d++) { flo[d] = -(double)ntess; fhi[d] = (double)ntess + 1.0; }
  if (sv->lat_bin) {
    for (int d = 0; d < 3; d++) { flo[d] = std::max(flo[d], sv->wlo[d]); fhi[d] = std::min(fhi[d], sv->whi[d]); }
  } else if (sv->window_on) {
    flo[0] = std::max(flo[0], sv->win_lo); fhi[0] = std::min(fhi[0], sv->win_hi);
  }
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int k = 0; k < 8; k++) {
    double fa = (k & 1) ? fhi[0] : flo[0], fb = (k & 2) ? fhi[1] : flo[1], fc = (k & 4) ? fhi[2] : flo[2];
    for (int d = 0; d < 3; d++) {
      double v = fa * L[d] + fb * L[3 + d] + fc * L[6 + d];
      mn[d] = std::min(mn[d], v);
      mx[d] = std::max(mx[d], v);
    }
  }
  GridParams g;
  double cell = (c->params.r_Rc + c->skin) * (1.0 + 1e-6);     // neighbour rows reach out to the cutoff + skin
  int gx = 1, gy = 1, gz = 1;
  const int zdiv = sv->lat_bin ? 4 : 1;     // z bins per cell edge (GridParams)
  for (int d = 0; d < 3; d++) { double pad = 1e-6 * (1.0 + fabs(mn[d]) + fabs(mx[d])); mn[d] -= pad; mx[d] += pad; }
  for (int it = 0; it < 200; it++) {
    gx = (int)floor((mx[0] - mn[0]) / cell) + 1;
    gy = (int)floor((mx[1] - mn[1]) / cell) + 1;
    gz = ((int)floor((mx[2] - mn[2]) / cell) + 1) * zdiv;
    if ((double)gx * gy * gz <= (double)sv->ncells_cap) break;
    cell *= 1.2599210498948732;
  }
  g.ox = mn[0]; g.oy = mn[1]; g.oz = mn[2];
  g.cell = cell; g.inv_cell = 1.0 / cell;
  g.zcell = cell / zdiv; g.inv_zcell = (double)zdiv / cell; g.zdiv = zdiv;
  g.gx = gx; g.gy = gy; g.gz = gz;
  g.ncell_mol = gx * gy * gz;
  g.ncells = g.ncell_mol;
  sv->hgrid = g;
  sv->grid_host = 1;
}

// Lattice path, windowed binning (tm_launch_lattice_bin, tm_nlist.cu): nothing is tessellated here; the view carries the
// lattice, its inverse and the fractional window per axis -- the cell (or this slab rank's share of it along the first
// lattice vector) widened by the largest interaction range, measured in plane spacings -- plus the host-laid cell grid.
static int prepare_lattice(tm_ctx* c, const double* xyz_dev, const int32_t* Z_dev, int64_t nreal, const double* lattice, int ntess, SysView* sv,
                           int rank = 0, int world = 1) {
  if (ntess < 1 || ntess > 8) { tm_set_error("ntess must be 1..8"); return TM_EINVAL; }
  const double* L = lattice;
  double det = L[0] * (L[4] * L[8] - L[5] * L[7]) - L[1] * (L[3] * L[8] - L[5] * L[6]) + L[2] * (L[3] * L[7] - L[4] * L[6]);
  if (fabs(det) < 1e-12) { tm_set_error("singular lattice"); return TM_EINVAL; }
  int64_t nslots = tess_count(nreal, ntess);
  *sv = make_view(c, nslots, 1, nslots, nreal, 1, nreal);
  double* gi = sv->ginv;   // inv(L), row-major: x = f L  =>  f_d = sum_r x_r inv[r][d]
  gi[0] = (L[4] * L[8] - L[5] * L[7]) / det; gi[1] = -(L[1] * L[8] - L[2] * L[7]) / det; gi[2] = (L[1] * L[5] - L[2] * L[4]) / det;
  gi[3] = -(L[3] * L[8] - L[5] * L[6]) / det; gi[4] = (L[0] * L[8] - L[2] * L[6]) / det; gi[5] = -(L[0] * L[5] - L[2] * L[3]) / det;
  gi[6] = (L[3] * L[7] - L[4] * L[6]) / det; gi[7] = -(L[0] * L[7] - L[1] * L[6]) / det; gi[8] = (L[0] * L[4] - L[1] * L[3]) / det;
  sv->slab_g[0] = gi[0]; sv->slab_g[1] = gi[3]; sv->slab_g[2] = gi[6];
  const double range = std::max(c->params.ee_cutoff_off, c->params.r_Rc) + c->skin + 0.05;
  for (int d = 0; d < 3; d++) {
    double gn = sqrt(gi[d] * gi[d] + gi[3 + d] * gi[3 + d] + gi[6 + d] * gi[6 + d]);   // 1 / plane spacing along axis d
    double halo = range * gn + 1e-5;    // real atoms are accepted up to 1e-6 outside [0, 1)
    sv->wlo[d] = -halo; sv->whi[d] = 1.0 + halo;
    if (d == 0 && world > 1) {
      sv->wlo[0] = (rank == 0) ? -halo - 1e-3 : (double)rank / world - halo;
      sv->whi[0] = (rank == world - 1) ? 1.0 + halo + 1e-3 : (double)(rank + 1) / world + halo;
    }
  }
  memcpy(sv->lat.v, lattice, 72);
  sv->lat.v[9] = 1.0 / (double)nreal;
  sv->lat_bin = 1; sv->lat_ntess = ntess;
  sv->xyz_real = xyz_dev; sv->Z_real = Z_dev;
  host_grid(c, sv, lattice, ntess);
  return TM_OK;
}

static int eval_lattice_impl(tm_ctx* c, const double* xyz, const int32_t* Z, int64_t nreal, const double* lattice, int ntess, int flags,
                             tm_outputs* out, bool use_host_grid) {
  int rc;
  c->launches = 0;
  size_t bx = (size_t)nreal * 24, bz = (size_t)nreal * 4;
  if ((rc = tm_host_stage(c, bx + bz + 64))) return rc;
  char* hs = (char*)c->h_stage;
  memcpy(hs, xyz, bx);
  memcpy(hs + bx, Z, bz);
  if ((rc = tm_buf(c, c->b_acc, bx + bz + 64))) return rc;
  c->timings_final = false;
  cudaEventRecord(c->ev[0], c->stream);
  TM_CUDA(cudaMemcpyAsync(c->b_acc.p, hs, bx + bz, cudaMemcpyHostToDevice, c->stream));
  SysView s;
  // windowed binning needs coordinates wrapped into the cell; unwrapped input (device flag 8) is redone by the caller with
  // use_host_grid = false: full tessellation (k_tessellate) and a bounding-box grid, as the reference does it
  if (use_host_grid) rc = prepare_lattice(c, (const double*)c->b_acc.p, (const int32_t*)((char*)c->b_acc.p + bx), nreal, lattice, ntess, &s);
  else rc = prepare_lattice_full(c, (const double*)c->b_acc.p, (const int32_t*)((char*)c->b_acc.p + bx), nreal, lattice, ntess, &s);
  if (rc) return rc;
  OutLayout o = out_layout(1, nreal);
  if ((rc = run_all(c, s, flags, o))) return rc;
  rc = deliver(c, s, flags, o, out, nreal);   // charges of the real atoms only: the image blocks are copies
  c->last.n_centres = nreal;
  return rc;
}

// The same evaluation replayed from a CUDA graph (H2D copy of the staged input, tessellation, every kernel of the step,
// D2H copy of the packed outputs and of the flag word).  Captured on the third consecutive call with the same shape;
// any (re)allocation, parameter / weight / stream change, or a different lattice or atom count falls back to the eager
// path and restarts the count.  Returns TM_OK, an error, or +1 = "use the eager path" (nothing was delivered).
static int lattice_graph_call(tm_ctx* c, const double* xyz, const int32_t* Z, int64_t nreal, const double* lattice, int ntess, int flags,
                              tm_outputs* out) {
  tm_ctx::LatGraph& G = c->lg;
  const int mask = out_mask(flags, out);
  // page-locked caller arrays are wired into the graph's copy nodes (no staging memcpy); the graph is then tied to
  // those addresses, and a call with other arrays re-captures like any other change of shape
  const void* pin[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  if (host_pinned(xyz) && host_pinned(Z)) { pin[0] = xyz; pin[1] = Z; }
  if ((mask & 1) && host_pinned(out->Ebp_atom)) pin[2] = out->Ebp_atom;
  if ((mask & 2) && host_pinned(out->charge)) pin[3] = out->charge;
  if ((mask & 4) && host_pinned(out->gradient)) pin[4] = out->gradient;
  bool same = G.nreal == nreal && G.ntess == ntess && G.flags == flags && G.outmask == mask && G.cfg_gen == c->cfg_gen &&
              G.alloc_gen == c->alloc_gen && memcmp(G.lat, lattice, 72) == 0 && memcmp(G.pin, pin, sizeof(pin)) == 0;
  if (!same) {
    if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
    G.nreal = nreal; G.ntess = ntess; G.flags = flags; G.outmask = mask; G.cfg_gen = c->cfg_gen; G.alloc_gen = c->alloc_gen;
    memcpy(G.lat, lattice, 72);
    memcpy(G.pin, pin, sizeof(pin));
    G.streak = 1; G.failed = false;
    return 1;
  }
  if (G.failed) return 1;
  if (!G.exec && ++G.streak < 3) return 1;
  int rc;
  OutLayout o = out_layout(1, nreal);
  size_t bx = (size_t)nreal * 24, bz = (size_t)nreal * 4, bytes = (size_t)o.total * 8;
  if (c->h_cap < std::max(bx + bz, bytes) + 64 || c->b_acc.cap < bx + bz + 64) return 1;   // the eager calls size these
  char* hs = (char*)c->h_stage;
  if (!pin[0]) {
    memcpy(hs, xyz, bx);
    memcpy(hs + bx, Z, bz);
  }
  void* const* direct = (void* const*)(pin + 2);
  if (!G.exec) {
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); G.failed = true; return 1; }
    c->launches = 0;
    SysView s;
    if (pin[0])
      rc = (cudaMemcpyAsync(c->b_acc.p, xyz, bx, cudaMemcpyHostToDevice, c->stream) == cudaSuccess &&
            cudaMemcpyAsync((char*)c->b_acc.p + bx, Z, bz, cudaMemcpyHostToDevice, c->stream) == cudaSuccess) ? TM_OK : TM_ECUDA;
    else
      rc = (cudaMemcpyAsync(c->b_acc.p, hs, bx + bz, cudaMemcpyHostToDevice, c->stream) == cudaSuccess) ? TM_OK : TM_ECUDA;
    if (!rc) rc = prepare_lattice(c, (const double*)c->b_acc.p, (const int32_t*)((char*)c->b_acc.p + bx), nreal, lattice, ntess, &s);
    if (!rc) rc = run_all(c, s, flags, o);
    if (!rc) rc = copy_packed_d2h(c, o, mask, direct);
    if (!rc && cudaMemcpyAsync(hs + bytes, c->b_flags.p, 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) rc = TM_ECUDA;
    cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
    if (rc || ce != cudaSuccess || !graph || c->alloc_gen != G.alloc_gen) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      G.failed = true;
      G.alloc_gen = c->alloc_gen;
      return 1;
    }
    ce = cudaGraphInstantiate(&G.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { cudaGetLastError(); G.exec = nullptr; G.failed = true; return 1; }
    G.launches = c->launches;
  }
  cudaEventRecord(c->ev[9], c->stream);
  TM_CUDA(cudaGraphLaunch(G.exec, c->stream));
  cudaEventRecord(c->ev[10], c->stream);
  TM_CUDA(cudaStreamSynchronize(c->stream));
  int32_t f0 = *(const int32_t*)(hs + bytes);
  c->last_flags = f0;
  if (f0 & 8) return 1;   // unwrapped input: the eager path redoes it with the bounding-box grid
  if (f0 & 2) { tm_set_error("more than %d radial neighbours of one centre", TM_NB_STRIDE); return TM_ECAP; }
  if (f0 & 4) { tm_set_error("more than %d neighbours inside the angular cutoff of one centre", TM_ANG_CAP); return TM_ECAP; }
  if ((rc = copy_out(c, flags, o, out, nreal, bytes, 0, direct))) return rc;
  tm_timings& t = c->last;
  memset(&t, 0, sizeof(t));
  c->timings_final = true;
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c->ev[9], c->ev[10]);
  t.total = ms;
  t.n_slots = tess_count(nreal, ntess);
  t.n_centres = nreal;
  t.launches = G.launches;
  c->launches = G.launches;
  c->cur_nslots = 0;
  return TM_OK;
}

extern "C" int tm_eval_lattice(tm_ctx* c, const double* xyz, const int32_t* Z, int64_t nreal, const double* lattice, int ntess, int flags,
                               tm_outputs* out) {
  if (c) c->nl_ok = false;
  if (!c || !xyz || !Z || !lattice || !out || nreal < 1) { tm_set_error("tm_eval_lattice: bad argument"); return TM_EINVAL; }
  int rc;
  TM_CUDA(cudaSetDevice(c->device));
  if ((rc = check_weights(c))) return rc;
  if ((rc = validate_Z(c, Z, nreal))) return rc;
  c->last_flags = 0;
  if (c->graphs_on && !(flags & TM_F_DESCRIPTORS)) {
    rc = lattice_graph_call(c, xyz, Z, nreal, lattice, ntess, flags, out);
    if (rc <= 0) return rc;                     // delivered from the graph, or failed
    if (c->last_flags & 8) return eval_lattice_impl(c, xyz, Z, nreal, lattice, ntess, flags, out, false);
  }
  rc = eval_lattice_impl(c, xyz, Z, nreal, lattice, ntess, flags, out, true);
  c->lg.alloc_gen = c->alloc_gen;              // allocations made by this eager call are part of the key
  // input not wrapped into the cell: the host-laid grid does not cover it, redo with the bounding-box pass
  if (rc == TM_EINVAL && (c->last_flags & 8)) rc = eval_lattice_impl(c, xyz, Z, nreal, lattice, ntess, flags, out, false);
  return rc;
}

extern "C" int tm_eval_lattice_dev(tm_ctx* c, const double* xyz_dev, const int32_t* Z_dev, int64_t nreal, const double* lattice, int ntess,
                                   int flags, double* e_dev, double* grad_dev, double* charge_dev) {
  if (!c || !xyz_dev || !Z_dev || !lattice || nreal < 1) { tm_set_error("tm_eval_lattice_dev: bad argument"); return TM_EINVAL; }
  int rc;
  TM_CUDA(cudaSetDevice(c->device));
  if ((rc = check_weights(c))) return rc;
  c->launches = 0;
  c->timings_final = false;
  cudaEventRecord(c->ev[0], c->stream);
  SysView s;
  bool reuse = (flags & TM_F_REUSE_NLIST) != 0;
  if (reuse) {
    if (!c->nl_ok || !c->hp.skin_on || c->nl_view.nreal != nreal || c->nl_view.lat_ntess != ntess || memcmp(c->nl_view.lat.v, lattice, 72) != 0) {
      tm_set_error("TM_F_REUSE_NLIST needs a previous tm_eval_lattice_dev call on this context with the same cell, and tm_set_skin > 0");
      return TM_ESTATE;
    }
    s = c->nl_view;
    s.xyz_real = xyz_dev; s.Z_real = Z_dev;
  } else {
    if ((rc = prepare_lattice(c, xyz_dev, Z_dev, nreal, lattice, ntess, &s))) return rc;
    c->nl_view = s;
    c->nl_ok = true;
  }
  OutLayout o = out_layout(1, nreal);
  if ((rc = run_all(c, s, flags, o, reuse))) return rc;
  const double* out = (const double*)c->b_out.p;
  if (e_dev) TM_CUDA(cudaMemcpyAsync(e_dev, out, 4 * 8, cudaMemcpyDeviceToDevice, c->stream));
  if (grad_dev && (flags & TM_F_FORCE)) TM_CUDA(cudaMemcpyAsync(grad_dev, out + o.off_grad, (size_t)3 * nreal * 8, cudaMemcpyDeviceToDevice, c->stream));
  if (charge_dev) TM_CUDA(cudaMemcpyAsync(charge_dev, out + o.off_charge, (size_t)nreal * 8, cudaMemcpyDeviceToDevice, c->stream));
  cudaEventRecord(c->ev[8], c->stream);
  c->cur_nslots = s.nslots;
  return TM_OK;
}

extern "C" int tm_get_timings(tm_ctx* c, tm_timings* t) {
  if (!c || !t) return TM_EINVAL;
  TM_CUDA(cudaSetDevice(c->device));
  TM_CUDA(cudaStreamSynchronize(c->stream));
  if (!c->timings_final) {      // (a graph replay has no per-stage events: its totals were stored by the call itself)
    SysView s;
    memset(&s, 0, sizeof(s));
    s.nslots = c->cur_nslots ? c->cur_nslots : c->last.n_slots;
    int64_t nc = c->last.n_centres;
    finish_timings(c, s);
    c->last.n_centres = nc;
  }
  *t = c->last;
  return TM_OK;
}

// ------------------------------------------------------------------------------------------------ slab-partitioned phases
// One process per GPU; every rank holds all positions, evaluates the centres of its own slab, and the
// host (tensormol_b200/parallel.py, torch.distributed/NCCL) performs the three small all-reduces
// between the phases.  See include/tmolb200.h.
__global__ void k_owned_qraw(const float* __restrict__ y, const int32_t* __restrict__ rowslot, int64_t nrows, double* __restrict__ q) {
  TM_PDL_PROLOGUE;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x) {
    int s = rowslot[r];
    if (s >= 0) q[s] = (double)y[r];
  }
}
__global__ void k_slab_e(const double* __restrict__ molacc, double* __restrict__ e, int add_ecc) {
  TM_PDL_PROLOGUE;
  e[0] = 0.0;
  e[1] = molacc[1];
  e[2] = add_ecc ? molacc[2] : 0.0;
  e[3] = molacc[3];
  e[4] = molacc[5];
  e[5] = 0.0;
}
__global__ void k_slab_set_dedq(double* __restrict__ molacc, const double* __restrict__ e) {
  TM_PDL_PROLOGUE; molacc[5] = e[4]; }

// ---- peer-memory exchange (see include/tmolb200.h) ---------------------------------------------------------------------
// symmetric buffer layout: [ q_raw f64[nreal] | e partials f64[16][8] | force partials f32[world][3 nreal] | flags ]
// flags (uint32, 64 B apart): [0..2] arrival counters of the three exchanges, [4..6] the epochs this rank has waited for.
struct PeerPtrs { char* p[16]; };

static void p2p_layout(int world, int64_t nreal, int64_t* off_q, int64_t* off_e, int64_t* off_g, int64_t* off_flag, int64_t* total) {
  auto up = [](int64_t v) { return (v + 255) / 256 * 256; };
  int64_t o = 0;
  *off_q = o; o += up(nreal * 8);
  *off_e = o; o += up(16 * 8 * 8);
  *off_g = o; o += up((int64_t)world * ((3 * nreal + 3) / 4 * 4) * 4);   // per-rank stride padded to 16 bytes
  *off_flag = o; o += 1024;
  *total = o;
}

extern "C" int64_t tm_slab_p2p_bytes(int world, int64_t nreal) {
  int64_t a, b, c2, d, t;
  p2p_layout(world, nreal, &a, &b, &c2, &d, &t);
  return t;
}

static int p2p_preload(tm_ctx* c);
// The exchange runs as separate store / signal / wait / sum launches.  TM_P2P_MERGED=1 selects the variants that fold the
// signal into the producing kernel (last block) and the wait into the force sum: two launches fewer per exchange, but the
// system-wide fence in every block costs more than the launches do (same box, 2 GPUs: 0.260 vs 0.247 ms at 3,300 atoms
// per rank, 0.572 vs 0.555 ms at 12,000), so they are kept for reference only.
static bool p2p_split_kernels() {
  static int v = -1;
  if (v < 0) v = getenv("TM_P2P_MERGED") ? 0 : 1;
  return v != 0;
}

extern "C" int tm_slab_p2p_setup(tm_ctx* c, int world, int rank, int64_t nreal, void* const* peer_base) {
  if (!c) { tm_set_error("tm_slab_p2p_setup: null context"); return TM_EINVAL; }
  c->cfg_gen++;
  if (world <= 1 || !peer_base) { c->p2p = tm_ctx::P2P(); return TM_OK; }
  if (world > 16 || rank < 0 || rank >= world || nreal < 1) { tm_set_error("tm_slab_p2p_setup: bad argument"); return TM_EINVAL; }
  tm_ctx::P2P q;
  q.on = 1; q.world = world; q.rank = rank; q.nreal = nreal;
  for (int r = 0; r < world; r++) {
    if (!peer_base[r]) { tm_set_error("tm_slab_p2p_setup: null peer buffer %d", r); return TM_EINVAL; }
    q.base[r] = (char*)peer_base[r];
  }
  int64_t total;
  p2p_layout(world, nreal, &q.off_q, &q.off_e, &q.off_g, &q.off_flag, &total);
  c->p2p = q;
  return p2p_preload(c);
}

static PeerPtrs peer_ptrs(const tm_ctx* c, int64_t off) {
  PeerPtrs pp;
  for (int r = 0; r < 16; r++) pp.p[r] = (r < c->p2p.world) ? c->p2p.base[r] + off : nullptr;
  return pp;
}

// q_raw of the owned centres, stored into every peer's copy (each slot has exactly one owner: no reduction needed)
__global__ void k_owned_qraw_p2p(const float* __restrict__ y, const int32_t* __restrict__ rowslot, int64_t nrows, PeerPtrs q, int world) {
  TM_PDL_PROLOGUE;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x) {
    int s = rowslot[r];
    if (s < 0) continue;
    double v = (double)y[r];
    for (int p = 0; p < world; p++) ((double*)q.p[p])[s] = v;
  }
}
// this rank's energy partials into slot [rank] of every peer
__global__ void k_slab_e_p2p(const double* __restrict__ molacc, int add_ecc, PeerPtrs e, int world, int rank) {
  TM_PDL_PROLOGUE;
  int p = threadIdx.x;
  if (p >= world) return;
  double* d = (double*)e.p[p] + 8 * rank;
  d[0] = 0.0; d[1] = molacc[1]; d[2] = add_ecc ? molacc[2] : 0.0; d[3] = molacc[3]; d[4] = molacc[5]; d[5] = 0.0;
}
// this rank's force partial (fp32, every atom: zeros outside its slab + halo) into slot [rank] of every peer
__global__ void k_push_grad_p2p(const float* __restrict__ F, int64_t n3, int64_t stride, PeerPtrs g, int world, int rank) {
  TM_PDL_PROLOGUE;
  int64_t n4 = n3 / 4;     // 3*nreal floats, float4 body + scalar tail; stride = n3 rounded up to 4 floats
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n4; t += (int64_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(F)[t];
    for (int p = 0; p < world; p++) reinterpret_cast<float4*>((float*)g.p[p] + (int64_t)rank * stride)[t] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n3 & 3)) {
    int64_t t = n4 * 4 + threadIdx.x;
    for (int p = 0; p < world; p++) ((float*)g.p[p] + (int64_t)rank * stride)[t] = F[t];
  }
}
// all of this device's earlier stores are out: bump arrival counter `which` on every peer (one thread per peer)
__global__ void k_p2p_signal(PeerPtrs f, int world, int which) {
  TM_PDL_PROLOGUE;
  __threadfence_system();
  int p = threadIdx.x;
  if (p < world) atomicAdd_system((unsigned int*)(f.p[p]) + 16 * which, 1u);
}
// The exchange kernels below carry their own signal: a block that has issued its peer stores fences them system-wide and
// counts itself done; the block that finishes last bumps arrival counter `which` on every peer (k_p2p_signal's job without
// the extra launch).  `done` is a device counter that the last block resets.
__device__ __forceinline__ void p2p_signal_when_last(PeerPtrs f, int world, int which, int32_t* done) {
  __shared__ int s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(done, 1) == (int)gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    if ((int)threadIdx.x < world) atomicAdd_system((unsigned int*)(f.p[threadIdx.x]) + 16 * which, 1u);
    if (threadIdx.x == 0) *done = 0;
  }
}
__global__ void k_owned_qraw_p2p_sig(const float* __restrict__ y, const int32_t* __restrict__ rowslot, int64_t nrows, PeerPtrs q, PeerPtrs f, int world,
                                     int32_t* __restrict__ done) {
  TM_PDL_PROLOGUE;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x) {
    int s = rowslot[r];
    if (s < 0) continue;
    double v = (double)y[r];
    for (int p = 0; p < world; p++) ((double*)q.p[p])[s] = v;
  }
  p2p_signal_when_last(f, world, 0, done);
}
// this rank's energy partials into slot [rank] of every peer, then the signal (one warp: lane p serves peer p)
__global__ void k_slab_e_p2p_sig(const double* __restrict__ molacc, int add_ecc, PeerPtrs e, PeerPtrs f, int world, int rank) {
  TM_PDL_PROLOGUE;
  int p = threadIdx.x;
  if (p < world) {
    double* d = (double*)e.p[p] + 8 * rank;
    d[0] = 0.0; d[1] = molacc[1]; d[2] = add_ecc ? molacc[2] : 0.0; d[3] = molacc[3]; d[4] = molacc[5]; d[5] = 0.0;
    __threadfence_system();
    atomicAdd_system((unsigned int*)(f.p[p]) + 16 * 1, 1u);
  }
}
__global__ void k_push_grad_p2p_sig(const float* __restrict__ F, int64_t n3, int64_t stride, PeerPtrs g, PeerPtrs f, int world, int rank,
                                    int32_t* __restrict__ done) {
  TM_PDL_PROLOGUE;
  int64_t n4 = n3 / 4;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n4; t += (int64_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(F)[t];
    for (int p = 0; p < world; p++) reinterpret_cast<float4*>((float*)g.p[p] + (int64_t)rank * stride)[t] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n3 & 3)) {
    int64_t t = n4 * 4 + threadIdx.x;
    for (int p = 0; p < world; p++) ((float*)g.p[p] + (int64_t)rank * stride)[t] = F[t];
  }
  p2p_signal_when_last(f, world, 2, done);
}
// the force sum with the wait for exchange 2 in front of it: every block spins (bounded) until all ranks have signalled.
// Epoch = the one k_p2p_wait of exchange 0 has set for this step (stable until the next step's phase B).
__global__ void k_wait_sum_grad_p2p(char* flags, int world, int32_t* errflags, const float* __restrict__ gparts, int64_t n3, int64_t stride,
                                    double* __restrict__ out) {
  TM_PDL_PROLOGUE;
  if (threadIdx.x == 0) {
    volatile unsigned int* cnt = (volatile unsigned int*)flags + 16 * 2;
    const unsigned int target = *((volatile unsigned int*)flags + 16 * 4) * (unsigned int)world;
    long long t0 = clock64();
    while ((int)(*cnt - target) < 0) {
      if (clock64() - t0 > 16000000000ll) { atomicOr(errflags, 32); break; }
      __nanosleep(100);
    }
    __threadfence_system();
  }
  __syncthreads();
  const volatile float* gp = gparts;     // written by the peers: no cached / hoisted reads
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n3; t += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < world; r++) s += (double)gp[(int64_t)r * stride + t];
    out[t] = s;
  }
}
// wait until every rank has signalled the next epoch of exchange `which` (epoch kept on the device: graph replayable)
__global__ void k_p2p_wait(char* flags, int world, int which, int32_t* errflags) {
  TM_PDL_PROLOGUE;
  volatile unsigned int* cnt = (volatile unsigned int*)flags + 16 * which;
  unsigned int* epoch = (unsigned int*)flags + 16 * (4 + which);
  unsigned int target = (*epoch + 1u) * (unsigned int)world;
  long long t0 = clock64();
  while ((int)(*cnt - target) < 0) {
    if (clock64() - t0 > 16000000000ll) { atomicOr(errflags, 32); break; }   // ~8 s: a peer never arrived
    __nanosleep(200);
  }
  *epoch += 1u;
  __threadfence_system();
}
// sum of the energy partials -> molacc[5] (sum dE/dq over all ranks) and the reduced energies, after the wait for
// exchange `which` (same protocol as k_p2p_wait)
__global__ void k_p2p_wait_reduce_e(char* flags, int world, int which, int32_t* errflags, const double* eparts, double* __restrict__ molacc,
                                    double* __restrict__ e_out) {
  TM_PDL_PROLOGUE;
  volatile unsigned int* cnt = (volatile unsigned int*)flags + 16 * which;
  unsigned int* epoch = (unsigned int*)flags + 16 * (4 + which);
  unsigned int target = (*epoch + 1u) * (unsigned int)world;
  long long t0 = clock64();
  while ((int)(*cnt - target) < 0) {
    if (clock64() - t0 > 16000000000ll) { atomicOr(errflags, 32); break; }   // ~8 s: a peer never arrived
    __nanosleep(200);
  }
  *epoch += 1u;
  __threadfence_system();
  const volatile double* ep = (const volatile double*)eparts;   // written by the peers: no cached / hoisted reads
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (int r = 0; r < world; r++)
    for (int k = 0; k < 6; k++) s[k] += ep[8 * r + k];
  molacc[5] = s[4];
  s[0] = s[1] + s[2] + s[3];
  for (int k = 0; k < 6; k++) e_out[k] = s[k];
}
__global__ void k_sum_grad_p2p(const float* __restrict__ gparts, int world, int64_t n3, int64_t stride, double* __restrict__ out) {
  TM_PDL_PROLOGUE;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n3; t += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < world; r++) s += (double)gparts[(int64_t)r * stride + t];
    out[t] = s;
  }
}

static int p2p_preload(tm_ctx* c) {
  // load the exchange kernels now: with lazy module loading the FIRST launch of a kernel can wait for the device to
  // drain, and a device that is spinning in k_p2p_wait only drains when its peers make progress
  cudaFuncAttributes fa;
  TM_CUDA(cudaSetDevice(c->device));
  TM_CUDA(cudaFuncGetAttributes(&fa, k_owned_qraw_p2p));
  TM_CUDA(cudaFuncGetAttributes(&fa, k_slab_e_p2p));
  TM_CUDA(cudaFuncGetAttributes(&fa, k_push_grad_p2p));
  TM_CUDA(cudaFuncGetAttributes(&fa, k_p2p_signal));
  TM_CUDA(cudaFuncGetAttributes(&fa, k_p2p_wait));
  TM_CUDA(cudaFuncGetAttributes(&fa, k_p2p_wait_reduce_e));
  TM_CUDA(cudaFuncGetAttributes(&fa, k_sum_grad_p2p));
  TM_CUDA(cudaFuncGetAttributes(&fa, k_owned_qraw_p2p_sig));
  TM_CUDA(cudaFuncGetAttributes(&fa, k_slab_e_p2p_sig));
  TM_CUDA(cudaFuncGetAttributes(&fa, k_push_grad_p2p_sig));
  TM_CUDA(cudaFuncGetAttributes(&fa, k_wait_sum_grad_p2p));
  return TM_OK;
}

extern "C" int tm_slab_phase_a(tm_ctx* c, const double* xyz_dev, const int32_t* Z_dev, int64_t nreal, const double* lattice, int ntess,
                               int rank, int world, double* qraw_dev) {
  if (c) c->nl_ok = false;
  if (!c || !xyz_dev || !Z_dev || !lattice || (!qraw_dev && !c->p2p.on) || world < 1 || rank < 0 || rank >= world) { tm_set_error("tm_slab_phase_a: bad argument"); return TM_EINVAL; }
  int rc;
  TM_CUDA(cudaSetDevice(c->device));
  if ((rc = check_weights(c))) return rc;
  c->launches = 0;
  c->timings_final = false;
  cudaEventRecord(c->ev[0], c->stream);
  SysView s;
  // the slab window along the first lattice vector (the owned slab plus the largest interaction range on both sides) is
  // laid out by prepare_lattice; everything an owned centre can interact with lies inside, nothing else is binned
  if ((rc = prepare_lattice(c, xyz_dev, Z_dev, nreal, lattice, ntess, &s, rank, world))) return rc;
  s.slab_rank = rank; s.slab_world = world; s.slab_api = 1;
  // a slab holds ~nreal/world centres; keep head-room for density fluctuations without a host round trip (device flag 64)
  if (world > 1) {
    s.ncent_max = std::min<int64_t>(nreal, nreal / world + nreal / (2 * world) + 4096);
    s.nrows = s.ncent_max + (int64_t)TM_ROW_TILE * c->hp.n_ele;
  }
  c->slab_view = s;
  if ((rc = stage_a(c, s))) return rc;
  // peer-memory exchange (the three phases are one stream of device work, possibly one CUDA graph): the backward GEMMs run on
  // the side stream while this stream pushes q_raw, waits for the peers and runs the pair kernel; joined in phase C.
  // Host collectives between the phases (NCCL fallback: one graph per phase) need the join inside this phase.
  if (c->p2p.on && overlap_backward(c, s)) {
    if ((rc = fork_backward(c, s))) return rc;
  } else if ((rc = tm_launch_mlp_backward(c, s))) {
    return rc;
  }
  cudaEventRecord(c->ev[6], c->stream);
  if (c->p2p.on) {
    if (c->p2p.world != world || c->p2p.rank != rank || c->p2p.nreal != nreal) { tm_set_error("tm_slab_phase_a: does not match tm_slab_p2p_setup"); return TM_EINVAL; }
    if ((rc = tm_buf(c, c->b_p2pdone, 64))) return rc;
    if (p2p_split_kernels()) {
      TM_LAUNCH(k_owned_qraw_p2p, nblk(s.nrows), 256, 0, c->stream, (const float*)c->b_y[TM_NET_CHARGE].p, (const int32_t*)c->b_rowslot.p, s.nrows,
                peer_ptrs(c, c->p2p.off_q), world);
      TM_LAUNCH(k_p2p_signal, 1, 32, 0, c->stream, peer_ptrs(c, c->p2p.off_flag), world, 0);
      c->launches += 2;
    } else {
      TM_LAUNCH(k_owned_qraw_p2p_sig, nblk(s.nrows), 256, 0, c->stream, (const float*)c->b_y[TM_NET_CHARGE].p, (const int32_t*)c->b_rowslot.p, s.nrows,
                peer_ptrs(c, c->p2p.off_q), peer_ptrs(c, c->p2p.off_flag), world, (int32_t*)c->b_p2pdone.p);
      c->launches += 1;
    }
  } else {
    TM_CUDA(cudaMemsetAsync(qraw_dev, 0, (size_t)nreal * 8, c->stream));
    TM_LAUNCH(k_owned_qraw, nblk(s.nrows), 256, 0, c->stream, (const float*)c->b_y[TM_NET_CHARGE].p, (const int32_t*)c->b_rowslot.p, s.nrows, qraw_dev);
    c->launches++;
  }
  TM_CUDA(cudaGetLastError());
  return TM_OK;
}

extern "C" int tm_slab_phase_b(tm_ctx* c, const double* qraw_dev, double* e_dev) {
  if (!c || ((!qraw_dev || !e_dev) && !c->p2p.on)) { tm_set_error("tm_slab_phase_b: bad argument"); return TM_EINVAL; }
  int rc;
  TM_CUDA(cudaSetDevice(c->device));
  SysView s = c->slab_view;
  int64_t nq = s.nreal;
  if ((rc = tm_buf(c, c->b_q, (size_t)nq * 8 * 2))) return rc;
  if (c->p2p.on) {   // every owner has stored its charges into this rank's copy once all ranks have signalled
    char* mine = c->p2p.base[c->p2p.rank];
    TM_LAUNCH(k_p2p_wait, 1, 1, 0, c->stream, mine + c->p2p.off_flag, c->p2p.world, 0, (int32_t*)c->b_flags.p);
    c->launches++;
    qraw_dev = (const double*)(mine + c->p2p.off_q);
  }
  TM_CUDA(cudaMemcpyAsync(c->b_q.p, qraw_dev, (size_t)nq * 8, cudaMemcpyDeviceToDevice, c->stream));
  SysView s2 = s;
  s2.slab_world = 2;   // any value > 1: tm_launch_charges then takes qraw from b_q instead of scattering its own rows
  if ((rc = tm_launch_charges(c, s2))) return rc;
  if ((rc = tm_launch_pair(c, s, TM_F_FORCE | TM_F_VDW))) return rc;
  if (!c->y_fused) {
    TM_LAUNCH(k_ebp, (int)((s.nrows + 255) / 256), 256, 0, c->stream, (const float*)c->b_y[TM_NET_ENERGY].p, (const int32_t*)c->b_rowslot.p, s.nrows, s.maxnatom,
                                                              (double*)c->b_molacc.p);
    c->launches++;
  }
  if (c->p2p.on) {
    if (p2p_split_kernels()) {
      TM_LAUNCH(k_slab_e_p2p, 1, 32, 0, c->stream, (const double*)c->b_molacc.p, c->hp.add_ecc, peer_ptrs(c, c->p2p.off_e), c->p2p.world, c->p2p.rank);
      TM_LAUNCH(k_p2p_signal, 1, 32, 0, c->stream, peer_ptrs(c, c->p2p.off_flag), c->p2p.world, 1);
      c->launches++;
    } else {
      TM_LAUNCH(k_slab_e_p2p_sig, 1, 32, 0, c->stream, (const double*)c->b_molacc.p, c->hp.add_ecc, peer_ptrs(c, c->p2p.off_e), peer_ptrs(c, c->p2p.off_flag),
                c->p2p.world, c->p2p.rank);
    }
  } else {
    TM_LAUNCH(k_slab_e, 1, 1, 0, c->stream, (const double*)c->b_molacc.p, e_dev, c->hp.add_ecc);
  }
  c->launches++;
  cudaEventRecord(c->ev[5], c->stream);
  TM_CUDA(cudaGetLastError());
  return TM_OK;
}

extern "C" int tm_slab_phase_c(tm_ctx* c, const double* e_dev, int flags, double* grad_dev) {
  if (!c || !e_dev || !grad_dev) { tm_set_error("tm_slab_phase_c: bad argument"); return TM_EINVAL; }
  int rc;
  TM_CUDA(cudaSetDevice(c->device));
  SysView s = c->slab_view;
  if (c->p2p.on) {
    char* mine = c->p2p.base[c->p2p.rank];
    TM_LAUNCH(k_p2p_wait_reduce_e, 1, 1, 0, c->stream, mine + c->p2p.off_flag, c->p2p.world, 1, (int32_t*)c->b_flags.p, (const double*)(mine + c->p2p.off_e),
                                                (double*)c->b_molacc.p, (double*)e_dev);
  } else {
    TM_LAUNCH(k_slab_set_dedq, 1, 1, 0, c->stream, (double*)c->b_molacc.p, e_dev);
  }
  c->launches++;
  if ((rc = join_backward(c))) return rc;
  if ((rc = tm_launch_force(c, s, flags))) return rc;
  if (c->p2p.on) {
    char* mine = c->p2p.base[c->p2p.rank];
    int64_t n3 = 3 * s.nreal, stride = (n3 + 3) / 4 * 4;
    if (p2p_split_kernels()) {
      TM_LAUNCH(k_push_grad_p2p, nblk(n3 / 4 + 1), 256, 0, c->stream, (const float*)c->b_F.p, n3, stride, peer_ptrs(c, c->p2p.off_g), c->p2p.world, c->p2p.rank);
      TM_LAUNCH(k_p2p_signal, 1, 32, 0, c->stream, peer_ptrs(c, c->p2p.off_flag), c->p2p.world, 2);
      TM_LAUNCH(k_p2p_wait, 1, 1, 0, c->stream, mine + c->p2p.off_flag, c->p2p.world, 2, (int32_t*)c->b_flags.p);
      TM_LAUNCH(k_sum_grad_p2p, nblk(n3), 256, 0, c->stream, (const float*)(mine + c->p2p.off_g), c->p2p.world, n3, stride, grad_dev);
      c->launches += 3;
    } else {
      TM_LAUNCH(k_push_grad_p2p_sig, nblk(n3 / 4 + 1), 256, 0, c->stream, (const float*)c->b_F.p, n3, stride, peer_ptrs(c, c->p2p.off_g), peer_ptrs(c, c->p2p.off_flag),
                c->p2p.world, c->p2p.rank, (int32_t*)c->b_p2pdone.p + 8);
      TM_LAUNCH(k_wait_sum_grad_p2p, nblk(n3), 256, 0, c->stream, mine + c->p2p.off_flag, c->p2p.world, (int32_t*)c->b_flags.p, (const float*)(mine + c->p2p.off_g),
                n3, stride, grad_dev);
      c->launches += 1;
    }
  } else {
    TM_LAUNCH(k_f2d, nblk(3 * s.nreal), 256, 0, c->stream, (const float*)c->b_F.p, grad_dev, 3 * s.nreal);
  }
  c->launches++;
  cudaEventRecord(c->ev[7], c->stream);
  cudaEventRecord(c->ev[8], c->stream);
  c->cur_nslots = s.nslots;
  TM_CUDA(cudaGetLastError());
  return TM_OK;
}
