// Internal declarations shared by the translation units of libtmolb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include "../../include/tmolb200.h"

#define TM_ROW_TILE 128          // GEMM row tile; element row groups are padded to this
#define TM_ANG_CAP 64            // angular neighbours per centre held in shared memory
#define TM_NB_STRIDE 256         // radial neighbour slots per centre row (liquid water: 48 max)
#define TM_MAX_ELEP (TM_MAX_ELE * (TM_MAX_ELE + 1) / 2)
// selu constants of the reference (TFInstance.py:365-369)
#define TM_SELU_ALPHA 1.6732632423543772848170429916717f
#define TM_SELU_SCALE 1.0507009873554804934193349852946f
#define TM_MAX_SYM 16            // max num_a_As / num_a_Rs
#define TM_BOHRPERA 1.889725989  // PhysicalData.py:30

void tm_set_error(const char* fmt, ...);
#define TM_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      tm_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return TM_ECUDA;                                                                        \
    }                                                                                         \
  } while (0)

// One atom of the cell-sorted copy: 32 bytes so a lane fetches it with two 128-bit loads.
struct __align__(32) SAtom {
  double x, y, z;
  int32_t slot;   // original slot index (mol*maxnatom + a, or image slot)
  int32_t e;      // element index into eles, -1 never stored
};

// Grid geometry: produced ON THE DEVICE from the bounding box (no host round trip), or laid out by the host from the
// lattice in the periodic-lattice path.
// lattice rows (9 doubles) + 1/natom, passed to k_tessellate by value
struct LatArgs { double v[10]; };

struct GridParams {
  double ox, oy, oz;   // origin
  double inv_cell;     // 1/cell edge
  double cell;
  int gx, gy, gz;      // cells per molecule box; gz counts the (possibly finer) z bins
  int ncell_mol;       // gx*gy*gz
  int ncells;          // nmol*ncell_mol
  // z is the fastest cell index, so any z-range of a column is ONE contiguous run of the cell-sorted array whatever the
  // bin height: the lattice path uses zdiv bins per cell edge, which clips the runs of the pair / neighbour kernels to
  // their sphere four times tighter at no cost per candidate (the molecule path keeps zdiv = 1)
  double zcell, inv_zcell;
  int zdiv;
};

// Constant hyper-parameters as the kernels consume them (fp32 + a few f64).
struct DevParams {
  int n_ele, n_elep;
  int eles[TM_MAX_ELE];
  int nRs_r, nRs_a, nAs, nsym;    // nsym = nAs*nRs_a
  int D, Dp;                      // descriptor width and padded width (multiple of 32)
  float r_Rc, a_Rc, eta, zeta;
  float pi_over_rRc, pi_over_aRc; // 3.14159265359/Rc  (truncated pi, RawSymFunc.py:935,1738)
  float zeta_pref;                // 2^(1-zeta)
  int zeta_is8;
  float Rs_r[64];
  float Rs_a[TM_MAX_SYM];
  float cosA[TM_MAX_SYM], sinA[TM_MAX_SYM];
  int8_t pair_index[TM_MAX_ELE][TM_MAX_ELE];
  // electrostatics, all in Bohr
  float R_lr, R_sr, alpha_b;      // EECutoffOff*B, Elu_Width*B, DSFAlpha/B
  float Zc, ZoverR_plus_Y;        // erfc(a R_lr)/R_lr ; Zc/R_lr + Yc
  float elu_a, elu_shift;
  float poly_width_b;             // Poly_Width*B
  float inv_poly_width_b;
  float erfc_c[12];               // erfcx(x) ~ sum_k c_k u^k, u = (x - erfc_mid)*erfc_ihalf, on [alpha R_sr, alpha R_lr]
  float erfc_mid, erfc_ihalf, erfc_fit_err;
  float sqrtC6[TM_MAX_ELE], Rvdw[TM_MAX_ELE];
  // the same pair potentials with every unit factor folded in, as the pair kernel evaluates them (r in Angstrom):
  //   LR   ex = 2^(pk_cex r^2), u = pk_ua r + pk_ub, erfc(aR)/R = (sum pk_pc[k] u^k) ex / r,
  //        kappa = erfc(aR)/R + pk_ka r + pk_kb,  B dkappa/dR = pk_BZY - (erfc(aR)/R + pk_c2 ex) / r
  //   ELU  ex = 2^(pk_ea r + pk_eb), kappa = elu_a ex + pk_ec, B dkappa/dR = pk_belu ex
  //   vdW  f0 = pk_c6[i][j] / r^6, X = pk_rs12[i][j] / r^12 (= 6 x^-12), t = pk_ta r
  float pk_rsr2, pk_rlr2, pk_cex, pk_ua, pk_ub, pk_pc[12], pk_ka, pk_kb, pk_c2, pk_BZY, pk_ea, pk_eb, pk_ec, pk_belu, pk_ta;
  float pk_c6[TM_MAX_ELE][TM_MAX_ELE], pk_rs12[TM_MAX_ELE][TM_MAX_ELE];
  int add_ecc;
  int activation;
  float act_alpha;
  double rr_exact, ra_exact;      // exact f64 cutoffs for the reference accept test
  int skin_on;                    // lists were built out to cutoff + skin: the kernels apply the cutoffs themselves
  float skin;
};

struct Layer {
  int K, N, Kp, Np;      // logical and padded dims
  float* W = nullptr;    // [Kp][Np] row-major fp32 (zero padded)
  float* WT = nullptr;   // [Np][Kp] (for the backward-data GEMM)
  float* b = nullptr;    // [Np]
  // split-fp16 planes for the tcgen05 path (x = hi + lo/2048): Ws = [W_hi | W_lo] each [Kp][Np]; WTs = [WT_hi | WT_lo] each [Np][Kp]
  uint16_t* Ws = nullptr;
  uint16_t* WTs = nullptr;
};

struct Net {
  Layer layers[TM_MAX_HIDDEN];   // hidden layers
  float* w_out = nullptr;        // [Hp_last]
  float b_out = 0.f;
  bool set[TM_MAX_HIDDEN + 1] = {false, false, false, false, false};
};

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

// Per-evaluation view passed to the launchers.
struct SysView {
  int64_t nslots;       // total slots (real + padding or images)
  int64_t nmol;
  int64_t maxnatom;     // slots per molecule
  int64_t nreal;        // periodic-images mode: centres are slots < nreal; else 0
  int periodic;         // 1 = images mode
  int slab_api;         // evaluated through tm_slab_phase_a/b/c (any world size): q_raw is combined over the ranks between the phases
  int64_t ncent_max;    // host upper bound on centre count
  int64_t nrows;        // ncent_max + TM_ROW_TILE*n_ele (allocated rows)
  int64_t ncells_cap;
  // slab ownership filter on centres (fraction of x-range), world==1 -> everything
  int slab_rank, slab_world;
  double slab_g[3];     // first row of the inverse lattice: frac = pos . slab_g
  // slab runs only: slots whose fractional coordinate lies outside [win_lo, win_hi] (slab + interaction halo) are not binned
  int window_on;
  double win_lo, win_hi;
  int win_ntess, win_ilo, win_ihi;   // image indices along the first lattice vector whose blocks can reach the window
  // lattice path: the cell grid is laid out on the host from the lattice (no bounding-box pass); grid_host = 1
  int grid_host;
  GridParams hgrid;
  // lattice path with windowed binning (tm_launch_lattice_bin): the images are never materialised as slots; every real
  // atom enumerates the lattice shifts that put it inside the fractional-coordinate window [wlo, whi] per axis (cell +
  // interaction halo, or a slab rank's share of it) and bins those, with the reference's slot ids and image arithmetic.
  int lat_bin;
  LatArgs lat;            // lattice rows + 1/natom
  double ginv[9];         // inverse lattice: frac_d = x ginv[d] + y ginv[3+d] + z ginv[6+d]
  double wlo[3], whi[3];
  int lat_ntess;
  const double* xyz_real;
  const int32_t* Z_real;
};

struct tm_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  // side stream for the nets' backward GEMMs of small problems (they depend on the forward pass only, not on the charge
  // exchange or the pair kernel): fork after the forward pass, join before the force kernel
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool fork_open = false;
  tm_model_desc desc;
  tm_params params;
  DevParams hp;                  // host copy
  DevParams* dp = nullptr;       // device copy
  Net nets[2][TM_MAX_ELE];
  int gemm_mode = TM_GEMM_TC_SPLIT;
  bool y_fused = false;   // the last forward pass already scattered q_raw and summed q_raw / Ebp per molecule (k_y_reduce)
  int last_flags = 0;                  // device flag word read by the last check_flags   // parity-preserving tensor-core path is the default
  int Hp[TM_MAX_HIDDEN];         // padded hidden widths
  int Hmax = 0;

  // ---- per-evaluation workspace (grow-only) ----
  DevBuf b_pos, b_Z, b_cellid, b_rank, b_count, b_cstart, b_sorted, b_satom, b_scan_tmp;
  DevBuf b_rowslot, b_rowsidx, b_rowofslot, b_blkcnt, b_rowmeta;
  DevBuf b_cntall, b_offall, b_pe, b_pairtab, b_lscan, b_gemm_ready[2], b_p2pdone;
  // Verlet skin (tm_set_skin): state of the last list-building lattice call
  double skin = 0.0;
  bool nl_ok = false;
  SysView nl_view;
  DevBuf b_pos0;
  // pair-potential tables of tm_pair.cu (rebuilt when the hyper-parameters change)
  uint64_t params_gen = 1, pairtab_gen = 0;
  int pt_kmin = 0, pt_nnodes = 0, pt_nfn = 0, pt_kink_k = -1, pt_kink_v = -1;
  DevBuf b_nbcnt, b_nboff, b_nbr, b_G, b_Gs, b_ypart, b_act[2][TM_MAX_HIDDEN], b_delta0, b_delta1, b_dG[2], b_y[2];
  DevBuf b_q, b_qs, b_dedq, b_F, b_acc, b_bbox, b_grid, b_flags, b_out, b_molacc;
  DevBuf b_natom;
  // host staging (pinned)
  void* h_stage = nullptr;
  size_t h_cap = 0;
  // results of compat calls (library-owned host memory)
  std::vector<int64_t> h_off, h_idx, h_rad, h_ang, h_milj, h_miljk;

  cudaEvent_t ev[12];
  bool ev_ok = false;
  tm_timings last;
  int launches = 0;

  // tm_eval_lattice replays its device work from a CUDA graph once the same call shape (atom count, lattice, flags,
  // configuration, buffer allocations) has been seen twice in a row; see lattice_graph_call in tm_api.cu
  uint64_t alloc_gen = 0;        // bumped by every (re)allocation of a workspace buffer
  uint64_t cfg_gen = 0;          // bumped by set_weights / set_params / set_gemm_mode / set_stream
  struct LatGraph {
    cudaGraphExec_t exec = nullptr;
    int64_t nreal = -1;
    int ntess = 0, flags = 0, streak = 0, launches = 0, outmask = 0;
    double lat[9];
    const void* pin[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // page-locked caller arrays wired into the copy nodes
    uint64_t alloc_gen = 0, cfg_gen = 0;
    bool failed = false;         // capture failed for this key: stay eager
  } lg;
  size_t gemm_ready_n[2] = {0, 0};
  void* tc_multi = nullptr;      // TcMulti scratch of the fused multi-layer GEMM launch
  void* tc_params = nullptr;     // TcParams scratch of tm_gemm_tc.cu (kernel argument block, one per context)
  void* tc_maps = nullptr;       // tensor-map cache of tm_gemm_tc.cu
  int graphs_on = 1;             // TM_NO_GRAPH=1 in the environment disables the replay
  bool timings_final = false;    // c->last already holds the timings of the last call (graph replay)

  // peer-memory exchange of the slab phases (tm_slab_p2p_setup)
  struct P2P {
    int on = 0, world = 1, rank = 0;
    int64_t nreal = 0;
    char* base[16] = {};          // every rank's symmetric buffer as addressable from this device
    int64_t off_q = 0, off_e = 0, off_g = 0, off_flag = 0;
  } p2p;

  // slab state
  int slab_rank = 0, slab_world = 1;
  int64_t cur_nslots = 0, cur_nreal = 0, cur_nmol = 0, cur_maxnatom = 0, cur_ncent = 0, cur_nrows = 0;
  int cur_periodic = 0;
  SysView slab_view;
  int slab_flags = 0;
};

// Programmatic dependent launch: every kernel of the step lets its successor be scheduled early (its blocks start as the
// SMs drain, launch latency and prologue overlap the predecessor's tail) and itself waits for its predecessors' memory
// before touching anything.  TM_NO_PDL=1 in the environment launches without the attribute (measurements).
int tm_pdl_enabled();
#ifdef __CUDACC__
#define TM_PDL_PROLOGUE asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory")
template <typename... KArgs, typename... Args>
static inline cudaError_t tm_launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = tm_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#define TM_LAUNCH(kern, grid, block, smem, stream, ...) tm_launch_kernel(kern, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)
__device__ __forceinline__ int cell_coord(double v, double o, double inv, int g) {
  int c = (int)floor((v - o) * inv);
  return c < 0 ? 0 : (c >= g ? g - 1 : c);
}
#endif

int tm_buf(tm_ctx* c, DevBuf& b, size_t bytes);
int tm_host_stage(tm_ctx* c, size_t bytes);

// ---- launchers (each returns TM_OK / error) ----
int tm_launch_tessellate(tm_ctx* c, const double* xyz_real, const int32_t* Z_real, int64_t nreal, const LatArgs& lattice, int ntess, int ilo, int ihi);
int tm_launch_nlist_build(tm_ctx* c, const SysView& s, double rc_grid);
int tm_launch_rows(tm_ctx* c, const SysView& s);
// centres this context is expected to hold for the view (a slab rank's share, not its row allocation)
static inline int64_t tm_expected_centres(const SysView& s) {
  return s.slab_world > 1 ? (s.periodic ? s.nreal : s.nslots) / s.slab_world : s.ncent_max;
}
int tm_launch_lattice_bin(tm_ctx* c, const SysView& s);
int tm_launch_lattice_refresh(tm_ctx* c, const SysView& s);   // Verlet-skin reuse: new positions into the existing cell-sorted records   // windowed binning + cell sort + centre rows of the lattice path
int tm_launch_neighbours(tm_ctx* c, const SysView& s);
int tm_launch_desc(tm_ctx* c, const SysView& s);
void tm_trace(tm_ctx* c, const char* what);   // TM_TRACE=1: synchronise + name the stage on stderr (tm_api.cu)
int tm_launch_mlp_forward(tm_ctx* c, const SysView& s);
int tm_launch_mlp_backward(tm_ctx* c, const SysView& s);
int tm_launch_charges(tm_ctx* c, const SysView& s);
int tm_launch_pair(tm_ctx* c, const SysView& s, int flags);
int tm_launch_force(tm_ctx* c, const SysView& s, int flags);

// generic CSR neighbour list for the MolEmb-compatible API
int tm_launch_nlist_csr(tm_ctx* c, const SysView& s, double rc, int do_perms, int64_t* total_out);

// GEMM back-ends (tm_gemm.cu): C[g] = act(A[g] * B[g] + bias) or the backward variant, grouped over row tiles
// fp32 mode: every pointer is float.  Tensor-core mode: A, B, Hmul and C (except the fp32 C of TM_EPI_NONE) address
// fp16 planes, the *2 members are the scaled "lo" planes, and B is given K-major as [N][K].
struct GemmGroup {
  const void* A;       // [rows][lda]
  const void* B;       // [K][ldb]  (K-major rows)
  const float* bias;   // [N] or nullptr
  const void* Hmul;    // backward: multiply result by act'(h) computed from Hmul [rows][ldc], or nullptr
  void* C;             // [rows][ldc]
  int lda, ldb, ldc, K, N;
  int ele;             // element whose row range this group covers
  const void* A2 = nullptr;
  const void* B2 = nullptr;
  const void* Hmul2 = nullptr;
  void* C2 = nullptr;
  int64_t rows_alloc = 0;
  const float* wout = nullptr;   // TM_EPI_ACT_OUT: output-layer weights [N] and partial sums [2*N/128][rows_alloc]
  float* ypart = nullptr;
};
void tm_gemm_tc_release(tm_ctx* c);
int tm_gemm_tc_launch_multi(tm_ctx* c, const GemmGroup* groups, int nlayers, int ngroups, const int* rowmeta_dev, int max_row_tiles,
                            int64_t expect_rows, const int* epilogues, bool backward);   // +1 = not applicable, launch layer by layer
int tm_launch_gemm(tm_ctx* c, const GemmGroup* groups, int ngroups, const int* rowmeta_dev, int max_row_tiles, int64_t expect_rows, int epilogue);
// TM_EPI_ACT_OUT (tensor-core mode, last hidden layer): h = act(z + b) is not stored; the epilogue emits the output layer's
// partial dot products  ypart[p][row] = sum_cols h*w_out  (p = 128-column half-tile index) and C = w_out * act'(h), the
// backward seed.
enum { TM_EPI_ACT = 0, TM_EPI_DACT = 1, TM_EPI_NONE = 2, TM_EPI_ACT_OUT = 3 };
