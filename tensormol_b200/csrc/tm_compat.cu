// MolEmb / Neighbors.py compatible entry points: the neighbour search runs on the device (K1, exact
// float64 accept test of C_API/MolEmb.cpp:1213-1218); the table layout of Neighbors.py:344-467 is
// assembled from the device CSR lists.
#include "tm_internal.h"
#include <algorithm>
#include <cstring>

__global__ void k_fill_const_i32(int32_t* p, int64_t n, int32_t v) {
  TM_PDL_PROLOGUE;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) p[t] = v;
}

static SysView compat_view(int64_t nslots, int64_t nmol, int64_t maxnatom, int64_t nreal, int periodic) {
  SysView s;
  memset(&s, 0, sizeof(s));
  s.nslots = nslots; s.nmol = nmol; s.maxnatom = maxnatom; s.nreal = nreal; s.periodic = periodic;
  s.ncent_max = periodic ? nreal : nslots;
  s.nrows = s.ncent_max;
  s.ncells_cap = nslots + 1024;
  s.slab_rank = 0; s.slab_world = 1; s.slab_api = 0;
  return s;
}

extern "C" int tm_nlist(tm_ctx* c, const double* xyz, int64_t n, int64_t nreal, double rc, int do_perms, const int64_t** offsets,
                        const int64_t** idx) {
  if (!c || !xyz || !offsets || !idx || n < 0 || nreal < 0 || nreal > n || !(rc > 0.0)) { tm_set_error("tm_nlist: bad argument"); return TM_EINVAL; }
  int r;
  TM_CUDA(cudaSetDevice(c->device));
  if (n == 0 || nreal == 0) {
    c->h_off.assign((size_t)nreal + 1, 0);
    c->h_idx.clear();
    c->h_idx.push_back(0);
    *offsets = c->h_off.data();
    *idx = c->h_idx.data();
    return TM_OK;
  }
  if ((r = tm_buf(c, c->b_pos, (size_t)n * 24))) return r;
  if ((r = tm_buf(c, c->b_Z, (size_t)n * 4))) return r;
  if ((r = tm_buf(c, c->b_flags, 64))) return r;
  TM_CUDA(cudaMemsetAsync(c->b_flags.p, 0, 64, c->stream));
  TM_CUDA(cudaMemcpyAsync(c->b_pos.p, xyz, (size_t)n * 24, cudaMemcpyHostToDevice, c->stream));
  int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  TM_LAUNCH(k_fill_const_i32, blocks, 256, 0, c->stream, (int32_t*)c->b_Z.p, n, 1);
  c->launches++;
  SysView s = compat_view(n, 1, n, nreal, nreal < n ? 1 : 0);
  if (nreal == n) { s.periodic = 0; s.nreal = 0; }
  int64_t total = 0;
  if ((r = tm_launch_nlist_csr(c, s, rc, do_perms, &total))) return r;
  if (c->h_idx.empty()) c->h_idx.push_back(0);
  *offsets = c->h_off.data();
  *idx = c->h_idx.data();
  return TM_OK;
}

// CSR over a padded set: rows = all slots (m*maxnatom + a); invalid slots have empty rows
static int set_csr(tm_ctx* c, int64_t nmol, int64_t maxnatom, double rc, std::vector<int64_t>& off, std::vector<int64_t>& idx) {
  SysView s = compat_view(nmol * maxnatom, nmol, maxnatom, 0, 0);
  int64_t total = 0;
  int r = tm_launch_nlist_csr(c, s, rc, 1, &total);
  if (r) return r;
  off = c->h_off;
  idx = c->h_idx;
  return TM_OK;
}

extern "C" int tm_pairs_triples_ele(tm_ctx* c, const double* xyzs, const int32_t* Zs, int64_t nmol, int64_t maxnatom, const int64_t* nnz,
                                    const int64_t* nreal, double rr, double ra, int64_t* Pn, int64_t* Tn, const int64_t** rad,
                                    const int64_t** ang, const int64_t** mil_j, const int64_t** mil_jk) {
  if (!c || !xyzs || !Zs || !nnz || !nreal || !Pn || !Tn || !rad || !ang || !mil_j || !mil_jk || nmol < 1 || maxnatom < 1) {
    tm_set_error("tm_pairs_triples_ele: bad argument");
    return TM_EINVAL;
  }
  int r;
  TM_CUDA(cudaSetDevice(c->device));
  int64_t nslots = nmol * maxnatom;
  std::vector<int32_t> hz((size_t)nslots);
  for (int64_t m = 0; m < nmol; m++)
    for (int64_t a = 0; a < maxnatom; a++) hz[m * maxnatom + a] = (a < nnz[m]) ? Zs[m * maxnatom + a] : 0;
  // element index per slot
  std::vector<int> ei((size_t)nslots, -1);
  for (int64_t t = 0; t < nslots; t++) {
    if (hz[t] <= 0) continue;
    for (int k = 0; k < c->desc.n_ele; k++)
      if (c->desc.eles[k] == hz[t]) ei[t] = k;
    if (ei[t] < 0) { tm_set_error("atomic number %d not in the model's element list", hz[t]); return TM_EINVAL; }
  }
  if ((r = tm_buf(c, c->b_pos, (size_t)nslots * 24))) return r;
  if ((r = tm_buf(c, c->b_Z, (size_t)nslots * 4))) return r;
  if ((r = tm_buf(c, c->b_flags, 64))) return r;
  TM_CUDA(cudaMemsetAsync(c->b_flags.p, 0, 64, c->stream));
  TM_CUDA(cudaMemcpyAsync(c->b_pos.p, xyzs, (size_t)nslots * 24, cudaMemcpyHostToDevice, c->stream));
  TM_CUDA(cudaMemcpyAsync(c->b_Z.p, hz.data(), (size_t)nslots * 4, cudaMemcpyHostToDevice, c->stream));
  TM_CUDA(cudaStreamSynchronize(c->stream));
  std::vector<int64_t> roff, ridx, aoff, aidx;
  if ((r = set_csr(c, nmol, maxnatom, rr, roff, ridx))) return r;
  if ((r = set_csr(c, nmol, maxnatom, ra, aoff, aidx))) return r;
  const int ne = c->desc.n_ele;
  auto pair_l = [&](int e1, int e2) { return (int)c->hp.pair_index[e1][e2]; };
  c->h_rad.clear(); c->h_ang.clear(); c->h_milj.clear(); c->h_miljk.clear();
  struct PR { int l; int64_t j; };
  struct TR { int l; int64_t k, j; };
  std::vector<PR> pr;
  std::vector<TR> tr;
  std::vector<int64_t> nb;
  for (int64_t m = 0; m < nmol; m++) {
    for (int64_t a = 0; a < nreal[m] && a < maxnatom; a++) {
      int64_t slot = m * maxnatom + a;
      if (hz[slot] <= 0) continue;
      // radial rows (mol,i,j,l) sorted by (l,j)    Neighbors.py:380
      pr.clear();
      for (int64_t t = roff[slot]; t < roff[slot + 1]; t++) {
        int64_t j = ridx[t];
        pr.push_back({ei[j], j - m * maxnatom});
      }
      std::sort(pr.begin(), pr.end(), [](const PR& x, const PR& y) { return x.l != y.l ? x.l < y.l : x.j < y.j; });
      int prev = -1, sl = 0;
      for (const PR& p : pr) {
        if (p.l != prev) { prev = p.l; sl = 0; }
        c->h_rad.insert(c->h_rad.end(), {m, a, p.j, (int64_t)p.l});
        c->h_milj.insert(c->h_milj.end(), {m, a, (int64_t)p.l, (int64_t)sl});
        sl++;
      }
      // triples: j<k by index, smaller atomic number first (Neighbors.py:173-180), sorted by (l,k,j) (:382)
      nb.clear();
      for (int64_t t = aoff[slot]; t < aoff[slot + 1]; t++) nb.push_back(aidx[t] - m * maxnatom);
      std::sort(nb.begin(), nb.end());
      tr.clear();
      for (size_t x = 0; x < nb.size(); x++)
        for (size_t y = x + 1; y < nb.size(); y++) {
          int64_t j = nb[x], k = nb[y];
          int zj = hz[m * maxnatom + j], zk = hz[m * maxnatom + k];
          if (zj > zk) std::swap(j, k);
          tr.push_back({pair_l(ei[m * maxnatom + j], ei[m * maxnatom + k]), k, j});
        }
      std::sort(tr.begin(), tr.end(), [](const TR& x, const TR& y) {
        if (x.l != y.l) return x.l < y.l;
        if (x.k != y.k) return x.k < y.k;
        return x.j < y.j;
      });
      prev = -1; sl = 0;
      for (const TR& t : tr) {
        if (t.l != prev) { prev = t.l; sl = 0; }
        c->h_ang.insert(c->h_ang.end(), {m, a, t.j, t.k, (int64_t)t.l});
        c->h_miljk.insert(c->h_miljk.end(), {m, a, (int64_t)t.l, (int64_t)sl});
        sl++;
      }
    }
  }
  (void)ne;
  *Pn = (int64_t)c->h_rad.size() / 4;
  *Tn = (int64_t)c->h_ang.size() / 5;
  if (c->h_rad.empty()) c->h_rad.push_back(0);
  if (c->h_ang.empty()) c->h_ang.push_back(0);
  if (c->h_milj.empty()) c->h_milj.push_back(0);
  if (c->h_miljk.empty()) c->h_miljk.push_back(0);
  *rad = c->h_rad.data();
  *ang = c->h_ang.data();
  *mil_j = c->h_milj.data();
  *mil_jk = c->h_miljk.data();
  return TM_OK;
}
