/*
 * tmolb200.h -- C-ABI of libtmolb200.so, the B200 (sm_100a) replacement for the native part
 * of TensorMol's per-step BP+EE energy/force evaluation.
 *
 * What it replaces in the reference (jparkhill/TensorMol, paths relative to the repo root):
 *   - the CPython extension `MolEmb` on the hot path:
 *       Make_NListNaive(xyz, rng, nreal, DoPerms)            C_API/MolEmb.cpp:1180-1247
 *       (method table C_API/MolEmb.cpp:2169-2230, parsed with "O!dii" at :1186)
 *   - the Python index assembly that consumes it:
 *       NeighborList.buildPairs / buildPairsAndTriples       TensorMol/ForceModifiers/Neighbors.py:75-201
 *       NeighborListSet.buildPairsAndTriplesWithEleIndex     Neighbors.py:344-423 (+Periodic :425-467)
 *       NeighborListSet.buildPairsWithBothEleIndex           Neighbors.py:323-342
 *   - the TensorFlow graph run by  Instances.evaluate / evaluate_periodic
 *       TensorMol/TFNetworks/TFMolInstanceDirect.py:5684-5711, 5918-5947
 *       (descriptors RawSymFunc.py:868-962,1696-1863,2223-2398; nets TFMolInstanceDirect.py:5164-5285,
 *        5774-5898; Coulomb/vdW RawSymFunc.py:1307-1465; tf.gradients :5761,5999)
 *
 * Conventions
 *   - plain C, no torch / numpy types.  All pointers are caller-owned unless stated.
 *   - every function returning int returns 0 on success, a negative TM_E* code on failure;
 *     tm_last_error() returns a thread-local message for the last failure on this thread.
 *   - one tm_ctx per host thread / GPU; calls on one ctx are serialised by the caller.
 *   - there is NO CPU fallback: every entry point needs a CUDA device and fails with
 *     TM_ECUDA otherwise.
 *   - "host" entry points take host buffers and include H2D/D2H copies; "_dev" entry points
 *     take device pointers (e.g. torch tensors' data_ptr()) and run on ctx's stream.
 */
#ifndef TMOLB200_H
#define TMOLB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TM_MAX_ELE 8
#define TM_MAX_HIDDEN 4

enum {
  TM_OK = 0,
  TM_EINVAL = -1,   /* bad argument */
  TM_ECUDA = -2,    /* CUDA runtime error / no device */
  TM_ESTATE = -3,   /* weights / params not set */
  TM_ECAP = -4      /* an internal capacity was exceeded (neighbour tile etc.) */
};

enum { TM_NET_CHARGE = 0, TM_NET_ENERGY = 1 };

/* activation ids: TFInstance.py:108-140 AssignActivation */
enum { TM_ACT_SIGMOID_WITH_PARAM = 0, TM_ACT_RELU = 1, TM_ACT_SOFTPLUS = 2, TM_ACT_TANH = 3, TM_ACT_SIGMOID = 4,
       TM_ACT_ELU = 5, TM_ACT_SELU = 6 };   /* PARAMS["NeuronType"] of TFInstance.AssignActivation (TFInstance.py:108-140); the
                                                 gaussian / square variants are not monotonic or not released and are refused */

/* GEMM arithmetic of the per-element MLPs */
enum {
  TM_GEMM_FP32 = 0,      /* fp32 FFMA tiles (parity reference mode of the library) */
  TM_GEMM_TC_SPLIT = 1,  /* tcgen05 kind::f16 on split operands x = hi + lo/2048 (two fp16 planes, 22 significant bits),
                            3 MMAs per K-step, fp32 accumulation in TMEM + registers (default) */
  TM_GEMM_TC_SPLIT_PAIR = 2, /* same arithmetic on CTA pairs: cta_group::2 MMAs of M = 256 over a cluster of two SMs,
                            each CTA stages its 128 A rows and half of the B rows (fewer L2 bytes per SM) */
  TM_GEMM_TC_SPLIT_N64 = 3, /* mode 1 with the 128 x 64 tile forced; mode 1 picks that tile by itself for launches with few
                            row tiles (slab ranks, small systems), this mode exists for tests and measurements */
  TM_GEMM_TC_SPLIT_N128 = 4 /* mode 1 with the 128 x 128 tile forced (tests and measurements) */
};

/* evaluation flags */
enum {
  TM_F_FORCE = 1,        /* compute the gradient / force */
  TM_F_VDW = 2,          /* HasVdw */
  TM_F_DESCRIPTORS = 4,  /* also copy descriptors out (parity tests) */
  TM_F_FOLD_IMAGES = 8,  /* NON-reference option: fold image-row gradients back onto their real atom */
  TM_F_REUSE_NLIST = 16  /* tm_eval_lattice_dev: keep the cell list / neighbour rows of the previous call (see tm_set_skin) */
};

typedef struct tm_ctx tm_ctx;

/* Network + element description.  eles ascending (TFMolInstanceDirect.py:1260-1267). */
typedef struct {
  int32_t n_ele;
  int32_t eles[TM_MAX_ELE];
  int32_t n_hidden;                 /* number of hidden layers, 1..TM_MAX_HIDDEN */
  int32_t hidden[TM_MAX_HIDDEN];    /* PARAMS["HiddenLayers"] */
} tm_model_desc;

/* Hyper-parameters (TMParams.py:26-38,150-165; SetANI1Param TFMolInstanceDirect.py:1293-1328). */
typedef struct {
  double r_Rc, a_Rc, eta, zeta;
  int32_t num_r_Rs, num_a_Rs, num_a_As;
  double ee_cutoff_on;    /* must be 0 (TFMolInstanceDirect.py:4366) */
  double ee_cutoff_off;   /* PARAMS["EECutoffOff"], Angstrom */
  double elu_width;       /* PARAMS["Elu_Width"] */
  double poly_width;      /* PARAMS["Poly_Width"] */
  double dsf_alpha;       /* PARAMS["DSFAlpha"] */
  double elu_shift, elu_alpha; /* DSF(), DSF_Gradient() of Util.py:172-192, computed by the host */
  int32_t add_ecc;        /* PARAMS["AddEcc"] */
  int32_t activation;     /* TM_ACT_* */
  double sigmoid_alpha;   /* PARAMS["sigmoid_alpha"] */
  double C6[TM_MAX_ELE];  /* a.u., TFMolInstanceDirect.py:3763-3767 */
  double Rvdw[TM_MAX_ELE];
} tm_params;

/* Outputs of one evaluation; any pointer may be NULL (skipped).  Shapes follow
 * TFMolManage.EvalBPDirectEEUpdateSet (TFMolManage.py:1260-1288): */
typedef struct {
  double* Etotal;     /* [nmol] Hartree */
  double* Ebp;        /* [nmol] */
  double* Ebp_atom;   /* [nmol*maxnatom] (periodic: [nreal]) */
  double* Ecc;        /* [nmol] */
  double* Evdw;       /* [nmol] */
  double* dipole;     /* [nmol*3] */
  double* charge;     /* [nmol*maxnatom] (periodic: [ntess_atoms], tiled like TFMolInstanceDirect.py:5892) */
  double* gradient;   /* dE/dx, Hartree/Angstrom, [nmol*maxnatom*3] (periodic: [nreal*3]) */
  float* descriptors; /* [nrows*D] in (mol,atom) row-major order, only with TM_F_DESCRIPTORS */
} tm_outputs;

/* Timings (ms, CUDA events on ctx's stream) of the last tm_eval*, per stage. */
typedef struct {
  float total, h2d, nlist, desc, mlp_fwd, pair, mlp_bwd, force, d2h;
  int64_t n_centres, n_slots, n_rad_pairs, n_ang_neigh, n_triples;
  int32_t launches;   /* kernels of this library launched by the call */
} tm_timings;

/* ---- lifecycle ---------------------------------------------------------------------- */
int tm_version(void);
const char* tm_last_error(void);
int tm_device_count(void);
tm_ctx* tm_create(int device, const tm_model_desc* desc, const tm_params* params);
void tm_destroy(tm_ctx* ctx);
int tm_set_params(tm_ctx* ctx, const tm_params* params);
/* y = a(x W + b):  W row-major [rows=fan_in][cols=fan_out] float64, b [cols].
 * layer in [0, n_hidden]; layer n_hidden is the linear regression layer (cols == 1).
 * Replaces tf.train.Saver.restore (TFMolInstanceDirect.py:5765). */
int tm_set_weights(tm_ctx* ctx, int net, int ele_index, int layer, const double* W, const double* b, int rows, int cols);
int tm_set_gemm_mode(tm_ctx* ctx, int mode);
int tm_get_gemm_mode(tm_ctx* ctx);
/* use an external CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); 0 = ctx-owned stream */
int tm_set_stream(tm_ctx* ctx, void* cuda_stream);
int tm_descriptor_width(tm_ctx* ctx);

/* ---- neighbour lists (replaces MolEmb.Make_NListNaive, C_API/MolEmb.cpp:1180-1247) --- */
/* xyz [n*3] float64 host.  Accept test is the reference's: sqrt(dx*dx+dy*dy+dz*dz)+1e-13 < rc.
 * Rows exist for i < nreal; do_perms as in the reference.  On return *offsets ([nreal+1]) and
 * *idx point to library-owned host memory, valid until the next call on ctx.  Each row is
 * sorted ascending (the reference's order is its sweep order; callers use sets). */
int tm_nlist(tm_ctx* ctx, const double* xyz, int64_t n, int64_t nreal, double rc, int do_perms,
             const int64_t** offsets, const int64_t** idx);

/* Pair/triple tables with element channels for a padded set (Neighbors.py:344-467).
 * xyzs [nmol*maxnatom*3], Zs [nmol*maxnatom] (0 = padding), nnz[nmol] atoms present,
 * nreal[nmol] centres (== nnz unless images).  Outputs are library-owned int64 host arrays:
 *   rad  [P*4] rows (mol,i,j,l) sorted by (mol,i,l,j)
 *   ang  [T*5] rows (mol,i,j,k,l) sorted by (mol,i,l,k,j), Z_j<=Z_k, j<k when Z equal
 *   mil_j [P*4] (mol,i,l,slot)   mil_jk [T*4] (mol,i,l,slot) */
int tm_pairs_triples_ele(tm_ctx* ctx, const double* xyzs, const int32_t* Zs, int64_t nmol, int64_t maxnatom,
                         const int64_t* nnz, const int64_t* nreal, double rr, double ra,
                         int64_t* P, int64_t* T, const int64_t** rad, const int64_t** ang,
                         const int64_t** mil_j, const int64_t** mil_jk);

/* ---- fused energy/force evaluation ---------------------------------------------------- */
/* Aperiodic set (EvalBPDirectEEUpdateSet/Single, TFMolManage.py:1260-1321). Host buffers. */
int tm_eval(tm_ctx* ctx, const double* xyzs, const int32_t* Zs, int64_t nmol, int64_t maxnatom,
            const int64_t* natom, int flags, tm_outputs* out);

/* Device-resident variant of tm_eval (the molecule-set form of tm_eval_lattice_dev; on-device MD / optimisation of
 * isolated molecules, SimpleMD.py:322-425): xyz_dev [nmol*maxnatom*3] f64 and Z_dev [nmol*maxnatom] i32 are DEVICE
 * pointers, Z zero in the padding slots (atoms of a molecule first); the outputs (any may be NULL) are device pointers:
 * e_dev [4*nmol] = Etotal | Ebp | Ecc | Evdw blocks (f64), grad_dev [nmol*maxnatom*3], charge_dev [nmol*maxnatom].
 * No host copy or synchronisation (capturable in a CUDA graph); atomic numbers are not validated (an element outside
 * the model's list is the caller's error); capacity flags surface from tm_sync. */
int tm_eval_dev(tm_ctx* ctx, const double* xyz_dev, const int32_t* Z_dev, int64_t nmol, int64_t maxnatom, int flags,
                double* e_dev, double* grad_dev, double* charge_dev);

/* Periodic, images supplied by the caller (the PeriodicForce callback form,
 * EvalBPDirectEEUpdateSinglePeriodic, TFMolManage.py:1323-1358): real atoms first, then
 * images with  slot = b*nreal + a  <->  real atom a  (Periodic.py:158-165). Host buffers. */
int tm_eval_images(tm_ctx* ctx, const double* xyz_tess, const int32_t* Z_tess, int64_t ntess_atoms, int64_t nreal,
                   int flags, tm_outputs* out);

/* Periodic from the primitive cell: wraps (Lattice.ModuloLattice must already be applied by
 * the caller), tessellates on the device exactly like Lattice.TessLattice (Periodic.py:131-168)
 * with the given ntess, then evaluates like tm_eval_images, except that out->charge is [nreal]
 * (the image blocks would be copies).  lattice [9] row vectors.
 * Repeated calls of one shape replay a captured CUDA graph (copies included).  Caller arrays in page-locked host
 * memory (cudaMallocHost / cudaHostRegister / a pinned torch tensor) — xyz and Z together, and each of Ebp_atom, charge,
 * gradient on its own — are wired into that graph's copy nodes and never pass through the library's staging buffer;
 * the graph is then tied to those addresses (other arrays: a few eager calls, then a new capture).  Pageable arrays
 * work the same way as before, through one staging memcpy each way. */
int tm_eval_lattice(tm_ctx* ctx, const double* xyz, const int32_t* Z, int64_t nreal, const double* lattice, int ntess,
                    int flags, tm_outputs* out);

/* Device-resident variant of tm_eval_lattice: xyz_dev [nreal*3] f64, Z_dev [nreal] i32 and the
 * outputs (any may be NULL) are DEVICE pointers: e_dev [4] = Etotal,Ebp,Ecc,Evdw (f64),
 * grad_dev [nreal*3] f64, charge_dev [nreal] f64.  No host synchronisation. */
int tm_eval_lattice_dev(tm_ctx* ctx, const double* xyz_dev, const int32_t* Z_dev, int64_t nreal, const double* lattice, int ntess,
                        int flags, double* e_dev, double* grad_dev, double* charge_dev);

/* Verlet skin (the idea at ForceModifiers/Periodic.py:224,262-268; the reference itself rebuilds every step).  With
 * skin > 0 the lattice path builds its cell list and neighbour rows out to cutoff + skin, and the descriptor, force and
 * pair kernels apply the cutoffs themselves, so every evaluation still sees exactly the neighbours inside the cutoffs.
 * A following tm_eval_lattice_dev call with TM_F_REUSE_NLIST then skips the neighbour build: it only refreshes the
 * positions (images with the reference's arithmetic).  Contract: same nreal / lattice / ntess, positions NOT re-wrapped
 * since the building call, no atom displaced by more than skin / 2 from where it was at the build (checked on the
 * device: tm_sync reports it).  skin = 0 (default) restores rebuild-every-call with no tests in the kernels. */
int tm_set_skin(tm_ctx* ctx, double skin_angstrom);

/* Slab-partitioned evaluation for multi-GPU runs: like tm_eval_lattice_dev but this rank only
 * evaluates centres whose row index r satisfies  lo <= r < hi  in the x-sorted centre order, in three
 * phases separated by the collectives the host performs (tensormol_b200/parallel.py):
 *   phase A: neighbour build, descriptors, both nets forward+unit-backward  -> qraw_dev [nreal] (0 for not-owned)
 *   (host: all-reduce qraw_dev)
 *   phase B: pair kernel for owned centres -> dedq_dev [nreal] (0 for not-owned), e_dev partial
 *   (host: all-reduce sum(dedq) inside e_dev[4])
 *   phase C: force kernel -> grad_dev [nreal*3] partial, e_dev[0..3] partial   (host: all-reduce) */
int tm_slab_phase_a(tm_ctx* ctx, const double* xyz_dev, const int32_t* Z_dev, int64_t nreal, const double* lattice, int ntess,
                    int rank, int world, double* qraw_dev);
int tm_slab_phase_b(tm_ctx* ctx, const double* qraw_dev, double* e_dev /* [6]: Etot,Ebp,Ecc,Evdw,sum_dedq,unused */);
int tm_slab_phase_c(tm_ctx* ctx, const double* e_dev, int flags, double* grad_dev);

/* Peer-memory exchange for the slab phases (one process per GPU of an NVLink/NVSwitch box; no reference counterpart).
 * Every rank owns a "symmetric" device buffer of tm_slab_p2p_bytes(world, nreal) bytes, ZERO-FILLED, that all peers can
 * address (CUDA IPC / torch symmetric memory); peer_base[r] is rank r's buffer as seen from this process.  Once set, the
 * three phases no longer need host collectives between them: the kernels that produce q_raw, the energy partials and
 * the force partials store them straight into every peer's buffer over NVLink, signal a per-peer flag, and the
 * consuming phase starts with a device-side wait on the local flag (bounded spin of about 8 s: a peer that never arrives raises
 * device flag 32, reported by tm_sync).  qraw_dev of phase A/B is then ignored, e_dev / grad_dev of phase C receive the
 * fully reduced energies [6] and dE/dx [nreal*3].  world <= 16.  tm_slab_p2p_setup(ctx, 0, ...) switches it off. */
int64_t tm_slab_p2p_bytes(int world, int64_t nreal);
int tm_slab_p2p_setup(tm_ctx* ctx, int world, int rank, int64_t nreal, void* const* peer_base);

int tm_get_timings(tm_ctx* ctx, tm_timings* t);
/* synchronise ctx's stream (for the _dev entry points) */
int tm_sync(tm_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* TMOLB200_H */
