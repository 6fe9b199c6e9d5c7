/*
 * TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's neighbour search, used by
 * tests/ (cross-check of oracle_np.py at sizes where numpy is slow) and by bench.py's CPU-baseline leg.
 * Nothing in the product path links or calls this file.
 *
 * Algorithm restated from C_API/MolEmb.cpp:1180-1247 (Make_NListNaive): argsort the atoms by x, then for
 * each atom sweep the following atoms in x-order until |dx| > rng; accept when
 *     sqrt(dx*dx + dy*dy + dz*dz) + 1e-13 < rng                      (MolEmb.cpp:1213-1218)
 * and at least one of the two indices is < nreal (:1208).  Row min(I,J) receives max(I,J); with DoPerms
 * the reverse entry is added when the larger index is < nreal too (:1220-1231).
 * Output is CSR (count pass, then fill pass) instead of Python lists.  Compile WITHOUT -ffast-math / FMA
 * contraction (the Makefile uses -O2 on x86-64, as the reference's setup.py does).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct { double x; int i; } xi_t;

static int cmp_x(const void* a, const void* b) {
  double xa = ((const xi_t*)a)->x, xb = ((const xi_t*)b)->x;
  return (xa > xb) - (xa < xb);
}

/* pass = 0: count into cnt[nreal]; pass = 1: fill idx using cursor[] (initialised to the row offsets) */
static void sweep(const double* xyz, const xi_t* y, int nat, int nreal, double rng, int do_perms, int pass, int64_t* cnt, int64_t* cursor, int32_t* idx) {
  for (int i = 0; i < nat; ++i) {
    int I = y[i].i;
    for (int j = i + 1; j < nat; ++j) {
      int J = y[j].i;
      if (!(I < nreal || J < nreal)) continue;
      if (fabs(xyz[I * 3] - xyz[J * 3]) > rng) break;
      double dx = xyz[I * 3 + 0] - xyz[J * 3 + 0];
      double dy = xyz[I * 3 + 1] - xyz[J * 3 + 1];
      double dz = xyz[I * 3 + 2] - xyz[J * 3 + 2];
      double dij = sqrt(dx * dx + dy * dy + dz * dz) + 0.0000000000001;
      if (dij < rng) {
        int lo = I < J ? I : J, hi = I < J ? J : I;
        if (pass == 0) {
          cnt[lo]++;
          if (hi < nreal && do_perms == 1) cnt[hi]++;
        } else {
          idx[cursor[lo]++] = hi;
          if (hi < nreal && do_perms == 1) idx[cursor[hi]++] = lo;
        }
      }
    }
  }
}

/* Returns the number of entries; *idx_out is malloc'ed (free with tm_oracle_free). off has nreal+1 entries. */
int64_t tm_oracle_nlist_naive(const double* xyz, int nat, int nreal, double rng, int do_perms, int64_t* off, int32_t** idx_out) {
  xi_t* y = (xi_t*)malloc(sizeof(xi_t) * (size_t)(nat > 0 ? nat : 1));
  for (int i = 0; i < nat; ++i) { y[i].x = xyz[i * 3]; y[i].i = i; }
  qsort(y, (size_t)nat, sizeof(xi_t), cmp_x);
  int64_t* cnt = (int64_t*)calloc((size_t)nreal + 1, sizeof(int64_t));
  sweep(xyz, y, nat, nreal, rng, do_perms, 0, cnt, 0, 0);
  off[0] = 0;
  for (int i = 0; i < nreal; ++i) off[i + 1] = off[i] + cnt[i];
  int64_t total = off[nreal];
  int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(total > 0 ? total : 1));
  for (int i = 0; i < nreal; ++i) cnt[i] = off[i];
  sweep(xyz, y, nat, nreal, rng, do_perms, 1, 0, cnt, idx);
  free(cnt);
  free(y);
  *idx_out = idx;
  return total;
}

void tm_oracle_free(void* p) { free(p); }
