// dpr1.cu -- dense-column handling: product-form LDL' of diag(d) + sum_k smult_k p_k p_k'
// (dpr1fact) and the product-form solves (fwdpr1, bwdpr1).
//
// Reference semantics:
//   dpr1fact.c:97-135    first pass over a column: pivot j is accepted while the implied multipliers
//                        stay below maxu, otherwise postponed (fi, d, t recurrences)
//   dpr1fact.c:224-240   second pass over the postponed rows, sorted by decreasing p^2
//   dpr1fact.c:280-477   dodpr1fact: dependent rows (d = 0), removal of one dependency when t > 0
//   dpr1fact.c:495-512   findnewdep after a subtraction (smult < 0, Lorentz trace columns)
//   dpr1fact.c:549-621   prodformfact: columns in order, each followed by a forward solve of the
//                        later columns that overlap it (auxfwdpr1.c:44-122)
//   fwdpr1.c:70-90, bwdpr1.c:65-162   product-form forward / backward solves
//
// These are scalar recurrences with data-dependent pivot decisions; only the work across the
// dense columns that follow (forward solves) and across right-hand sides is independent.  The kernels
// keep the decision chain on one thread of a CTA and spread the independent pieces over the rest:
// this keeps the factor on the device between blkchol and the solves, it is not a throughput kernel
// (SURVEY.md section 8d lists it as latency-bound; it only runs when getdense finds dense columns).
#include <algorithm>
#include "sb_internal.h"

namespace sb {

struct KD { double r; int k; };

// forward solve with L(p,beta) = I + tril(p beta', -1), natural order (auxfwdpr1.c:44-77)
__device__ void d_fwipr1(double *y, const double *p, const double *beta, int m, int n) {
  if (n < 1) return;
  double yi = y[0], betai = beta[0], t = 0.0;
  int i = 1;
  for (; i < n; i++) {
    t += yi * betai;
    yi = (y[i] -= t * p[i]);
    betai = beta[i];
  }
  if (n < m) {
    t += yi * betai;
    for (; i < m; i++) y[i] -= t * p[i];
  }
}
// ordered variant (auxfwdpr1.c:79-122)
__device__ void d_fwipr1o(double *y, const int *perm, const double *p, const double *beta, int m, int n) {
  if (n < 1) return;
  double yi = y[perm[0]], betai = beta[0], t = 0.0;
  int i = 1;
  for (; i < n; i++) {
    t += yi * betai;
    const int pi = perm[i];
    yi = (y[pi] -= t * p[pi]);
    betai = beta[i];
  }
  if (n < m) {
    t += yi * betai;
    for (; i < m; i++) { const int pi = perm[i]; y[pi] -= t * p[pi]; }
  }
}
// backward solves (bwdpr1.c:65-113)
__device__ void d_bwipr1(double *y, const double *p, const double *beta, int m, int n) {
  if (n < 1) return;
  double t = 0.0;
  for (int i = n; i < m; i++) t += p[i] * y[i];
  for (int i = n; i > 0; i--) {
    const double yi = (y[i - 1] -= t * beta[i - 1]);
    t += p[i - 1] * yi;
  }
}
__device__ void d_bwipr1o(double *y, const int *perm, const double *p, const double *beta, int m, int n) {
  if (n < 1) return;
  double t = 0.0;
  for (int i = m - 1; i >= n; i--) { const int pi = perm[i]; t += p[pi] * y[pi]; }
  for (int i = n; i > 0; i--) {
    const int pi = perm[i - 1];
    const double yi = (y[pi] -= t * beta[i - 1]);
    t += p[pi] * yi;
  }
}

// One rank-1 step: (D + smult p p')(perm) = L diag(d_new(perm)) L', L = I + tril(p(perm) beta', -1).
// Runs on ONE thread.  Returns 1 if rows were re-ordered (perm written), 0 for the natural order.
__device__ int d_rank1_factor(double *beta, int *perm, double *d, double smult, const double *p, int m, int *pn,
                              int *dep, int *pndep, double maxu, double *fi, double *mu, KD *kd) {
  if (smult == 0.0) { *pn = 0; return 0; }
  double t = 1.0 / smult;
  int ndep = *pndep;
  for (int i = 0; i < m; i++) fi[i] = p[i] * p[i];
  const double maxusq_scale = maxu;            // compared as (maxu*fij)^2 like the reference
  if (dep[0] >= m) {
    // ---- no dependent row among the first m: natural order first, postponed rows afterwards
    *pn = m;
    double h = 0.0;
    for (int i = m; i > 0; i--) { mu[i - 1] = h; h = fmax(h, fi[i - 1]); }
    int nph2 = 0;
    double muph2 = 0.0;
    for (int j = 0; j < m; j++) {
      const double dj = d[j], x = fi[j];
      const double fij = x + t * dj;
      const double lim = maxusq_scale * fij;
      if (x * fmax(muph2, mu[j]) <= lim * lim) { fi[j] = fij; d[j] = fij / t; t = fij / dj; }
      else { kd[nph2].r = x; kd[nph2].k = j; nph2++; muph2 = fmax(muph2, x); }
    }
    if (nph2 == 0) {
      for (int i = 0; i < m; i++) beta[i] = p[i] / fi[i];
      return 0;
    }
    int w = 0, q = 0;                          // accepted rows keep their order
    for (int j = 0; j < m; j++) {
      if (q < nph2 && kd[q].k == j) { q++; continue; }
      perm[w] = j; beta[w] = p[j] / fi[j]; w++;
    }
    // postponed rows by decreasing p^2 (insertion sort; ties keep ascending row order)
    for (int a = 1; a < nph2; a++) {
      KD key = kd[a]; int b = a - 1;
      while (b >= 0 && kd[b].r < key.r) { kd[b + 1] = kd[b]; b--; }
      kd[b + 1] = key;
    }
    for (int a = 0; a < nph2; a++) {
      const int j = kd[a].k;
      const double dj = d[j];
      const double fij = (kd[a].r += t * dj);
      d[j] = fij / t; t = fij / dj;
      perm[w + a] = j; beta[w + a] = p[j] / fij;
    }
    return 1;
  }
  // ---- some d(i) = 0 among the first m rows
  double psqrdep = 0.0;
  int jd = 0, i;
  for (i = 0; dep[i] < m; i++)
    if (fi[dep[i]] > psqrdep) { jd = i; psqrdep = fi[dep[i]]; }
  int idep, deldep = 0;
  double h;
  if (psqrdep > 0.0) {
    idep = dep[jd];
    if (t > 0.0) {
      deldep = 1;
      for (int a = jd; a < ndep; a++) dep[a] = dep[a + 1];      // shifts the tail entry too
      h = maxu * maxu * psqrdep;
      dep[ndep] = idep;
      *pndep = --ndep;
    } else { h = psqrdep; deldep = 0; }
  } else { idep = dep[0]; h = 0.0; deldep = 0; }
  int j = 0, back = m;
  for (i = 0; i < idep; i++) { if (fi[i] > h) perm[j++] = i; else perm[--back] = i; }
  for (++i; i < m; i++) { if (fi[i] > h) perm[j++] = i; else perm[--back] = i; }
  perm[j] = idep;
  int n = j;
  *pn = j + deldep;
  for (i = n; i > 0; i--) { mu[i - 1] = h; h = fmax(h, fi[perm[i - 1]]); }
  int nph2 = 0, jnz = 0;
  double muph2 = 0.0;
  for (i = 0; i < n; i++) {
    const int k = perm[i];
    const double dj = d[k], x = fi[k];
    const double fij = x + t * dj;
    const double lim = maxu * fij;
    if (x * fmax(muph2, mu[i]) <= lim * lim) { fi[k] = fij; perm[jnz++] = k; d[k] = fij / t; t = fij / dj; }
    else { kd[nph2].r = x; kd[nph2].k = k; nph2++; muph2 = fmax(muph2, x); }
  }
  n -= nph2;
  for (i = 0; i < n; i++) beta[i] = p[perm[i]] / fi[perm[i]];
  if (nph2) {
    for (int a = 1; a < nph2; a++) {
      KD key = kd[a]; int b = a - 1;
      while (b >= 0 && kd[b].r < key.r) { kd[b + 1] = kd[b]; b--; }
      kd[b + 1] = key;
    }
    for (int a = 0; a < nph2; a++) {
      const int k = kd[a].k;
      const double dj = d[k];
      const double fij = (kd[a].r += t * dj);
      d[k] = fij / t; t = fij / dj;
      perm[n + a] = k; beta[n + a] = p[k] / fij;
    }
  }
  if (deldep) { d[idep] = fi[idep] / t; beta[n + nph2] = 1.0 / p[idep]; }
  return 1;
}

// dep(0:ndep) sorted + previously removed dependencies behind it (dpr1fact.c:495-512)
__device__ int d_findnewdep(int *dep, int ndep, int maxndep, const double *d) {
  int i;
  for (i = ndep + 1; i <= maxndep; i++) if (d[dep[i]] <= 0.0) break;
  if (i > maxndep) return 0;
  const int idep = dep[i];
  int j = 0;
  while (j < ndep && dep[j] <= idep) j++;           // first j with dep[j] > idep
  for (int a = i; a > j; a--) dep[a] = dep[a - 1];
  dep[j] = idep;
  return 1;
}

// The whole product-form factorisation, one CTA.  xs[k] = number of rows of column k (dz.jc[k+1]);
// p holds the columns back to back (column k starts at sum_{j<=k-1} xs[j] ... see poff[]).
__global__ void __launch_bounds__(256)
prodform_kernel(int n, const int *xs, const long long *poff, double *p, int *pivperm, double *beta, int *betajc,
                double *d, int *ordered, const int *colperm, const int *firstpiv, const double *smult, int *dep,
                int *ndep_io, double maxu, double *fi, double *mu, KD *kd, long long *permoff_out) {
  __shared__ int s_useperm, s_nk, s_inz;
  __shared__ long long s_permoff;
  if (threadIdx.x == 0) { s_inz = 0; s_permoff = 0; }
  __syncthreads();
  const int maxndep = *ndep_io;
  for (int k = 0; k < n; k++) {
    const int colk = colperm[k];
    const int mk = xs[k];
    double *pk = p + poff[k];
    if (threadIdx.x == 0) {
      betajc[k] = s_inz;
      int nk = 0, ndep = *ndep_io;
      int up = d_rank1_factor(beta + s_inz, pivperm + s_permoff, d, smult[colk], pk, mk, &nk, dep, &ndep, maxu, fi, mu, kd);
      ordered[k] = up;
      if (smult[colk] < 0.0) ndep += d_findnewdep(dep, ndep, maxndep, d);
      *ndep_io = ndep;
      s_useperm = up; s_nk = nk;
    }
    __syncthreads();
    if (smult[colk] != 0.0) {
      const double *betak = beta + s_inz;
      const int *permk = pivperm + s_permoff;
      for (int j = k + 1 + threadIdx.x; j < n; j += blockDim.x) {
        if (firstpiv[colperm[j]] <= k) {
          double *pj = p + poff[j];
          if (s_useperm) d_fwipr1o(pj, permk, pk, betak, mk, s_nk);
          else d_fwipr1(pj, pk, betak, mk, s_nk);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (smult[colk] != 0.0 && s_useperm) s_permoff += mk;
      s_inz += s_nk;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { betajc[n] = s_inz; *permoff_out = s_permoff; }
}

// product-form solves: one thread per right-hand side (columns of y are independent)
__global__ void prodform_solve_kernel(int backward, int nrhs, int m, int nden, int dznnz, const int *dzir, const int *xs,
                                      const long long *poff, const long long *permoff, const double *p, const int *pivperm,
                                      const double *beta, const int *betajc, const int *ordered, double *y, double *fwork) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= nrhs) return;
  double *yy = y + (long long)col * m, *f = fwork + (long long)col * dznnz;
  for (int i = 0; i < dznnz; i++) f[i] = yy[dzir[i]];
  if (!backward) {
    for (int k = 0; k < nden; k++) {
      const int nk = betajc[k + 1] - betajc[k];
      if (ordered[k]) d_fwipr1o(f, pivperm + permoff[k], p + poff[k], beta + betajc[k], xs[k], nk);
      else d_fwipr1(f, p + poff[k], beta + betajc[k], xs[k], nk);
    }
  } else {
    for (int k = nden - 1; k >= 0; k--) {
      const int nk = betajc[k + 1] - betajc[k];
      if (ordered[k]) d_bwipr1o(f, pivperm + permoff[k], p + poff[k], beta + betajc[k], xs[k], nk);
      else d_bwipr1(f, p + poff[k], beta + betajc[k], xs[k], nk);
    }
  }
  for (int i = 0; i < dznnz; i++) yy[dzir[i]] = f[i];
}

}  // namespace sb
using namespace sb;

extern "C" {

// [Lden,d] = dpr1fact(x,d,Lsymb,smult,maxu)   (dpr1fact.c:630-848).
//   x: sparse m x n (CSC)  = L\Ad ; dz: Lsymb.dz pattern (jc[n+1], ir[dznnz]); colperm/firstpiv 0-based.
// Outputs: p (pnnz = sum_k dz.jc[k+1]), beta (<= pnnz, *nbeta), betajc (n+1, 0-based), pivperm
// (*npivperm entries, 0-based), dopiv (n), d (m, updated copy).
int sb200_dpr1fact(sb_idx m, sb_idx n, const sb_idx *xjc, const sb_idx *xir, const double *xpr, const double *d_in,
                   const sb_idx *dzjc, const sb_idx *dzir, const sb_idx *colperm, const sb_idx *firstpiv,
                   const double *smult, double maxu, double *p_out, double *beta_out, sb_idx *betajc_out,
                   sb_idx *pivperm_out, double *dopiv_out, double *d_out, sb_idx *nbeta, sb_idx *npivperm) {
  SB_TRY(ensure_init());
  const sb_idx dznnz = dzjc[n];
  SB_CHECK(dznnz <= m, "dpr1fact: Lsymb.dz has more rows than d");
  std::vector<int> xs(n), colp(n), firstp(n), dzir32((size_t)std::max<sb_idx>(dznnz, 1));
  std::vector<long long> poff(n + 1, 0);
  sb_idx pnnz = 0;
  for (sb_idx k = 0; k < n; k++) {
    xs[k] = (int)dzjc[k + 1]; poff[k] = pnnz; pnnz += dzjc[k + 1];
    SB_CHECK(colperm[k] >= 0 && colperm[k] < n, "dpr1fact: Lsymb.perm out of range");
    colp[k] = (int)colperm[k]; firstp[k] = (int)firstpiv[k];
  }
  poff[n] = pnnz;
  std::vector<sb_idx> invrow((size_t)std::max<sb_idx>(m, 1), -1);
  for (sb_idx i = 0; i < dznnz; i++) { SB_CHECK(dzir[i] >= 0 && dzir[i] < m, "dpr1fact: Lsymb.dz row out of range"); invrow[dzir[i]] = i; dzir32[i] = (int)dzir[i]; }
  // p(invrowperm,:) = x(:,colperm); d(1:dznnz) = lab(dz.ir); dep = find(d <= 0)   (dpr1fact.c:735-760)
  std::vector<double> p((size_t)std::max<sb_idx>(pnnz, 1), 0.0), dd((size_t)std::max<sb_idx>(dznnz, 1), 0.0);
  for (sb_idx j = 0; j < n; j++) {
    const sb_idx pj = colperm[j];
    for (sb_idx i = xjc[pj]; i < xjc[pj + 1]; i++) {
      const sb_idx r = invrow[xir[i]];
      SB_CHECK(r >= 0 && r < dzjc[j + 1], "dpr1fact: x has a nonzero outside Lsymb.dz");
      p[poff[j] + r] = xpr[i];
    }
  }
  std::vector<int> dep((size_t)m + 2, 0);
  int ndep = 0;
  for (sb_idx i = 0; i < dznnz; i++) { dd[i] = d_in[dzir[i]]; if (dd[i] <= 0.0) dep[ndep++] = (int)i; }
  dep[ndep] = (int)m;
  memcpy(d_out, d_in, sizeof(double) * m);
  *nbeta = 0; *npivperm = 0;
  for (sb_idx k = 0; k <= n; k++) betajc_out[k] = 0;
  if (n == 0) return 0;
  arena_reset();
  cudaStream_t st = ctx().stream;
  int *d_xs = arena<int>(n), *d_colp = arena<int>(n), *d_first = arena<int>(n), *d_pivperm = arena<int>(std::max<sb_idx>(pnnz, 1)),
      *d_betajc = arena<int>(n + 1), *d_ordered = arena<int>(n), *d_dep = arena<int>(m + 2), *d_ndep = arena<int>(1);
  long long *d_poff = arena<long long>(n + 1), *d_permoff = arena<long long>(1);
  double *d_p = arena<double>(std::max<sb_idx>(pnnz, 1)), *d_beta = arena<double>(std::max<sb_idx>(pnnz, 1)), *d_d = arena<double>(std::max<sb_idx>(dznnz, 1)),
         *d_smult = arena<double>(n), *d_fi = arena<double>(std::max<sb_idx>(dznnz, 1)), *d_mu = arena<double>(std::max<sb_idx>(dznnz, 1));
  KD *d_kd = arena<KD>(std::max<sb_idx>(dznnz, 1));
  SB_CHECK(d_xs && d_colp && d_first && d_pivperm && d_betajc && d_ordered && d_dep && d_ndep && d_poff && d_permoff && d_p && d_beta &&
           d_d && d_smult && d_fi && d_mu && d_kd, "dpr1fact: out of device memory");
  SB_CUDA(cudaMemcpyAsync(d_xs, xs.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_colp, colp.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_first, firstp.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_poff, poff.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_p, p.data(), sizeof(double) * p.size(), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_d, dd.data(), sizeof(double) * dd.size(), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_smult, smult, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_dep, dep.data(), sizeof(int) * (m + 2), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_ndep, &ndep, sizeof(int), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemsetAsync(d_beta, 0, sizeof(double) * std::max<sb_idx>(pnnz, 1), st));
  prodform_kernel<<<1, 256, 0, st>>>((int)n, d_xs, d_poff, d_p, d_pivperm, d_beta, d_betajc, d_d, d_ordered, d_colp, d_first,
                                     d_smult, d_dep, d_ndep, maxu, d_fi, d_mu, d_kd, d_permoff);
  SB_LAUNCH_CHECK_N("prodform_kernel");
  std::vector<int> h_betajc(n + 1), h_ordered(n), h_pivperm((size_t)std::max<sb_idx>(pnnz, 1));
  std::vector<double> h_beta((size_t)std::max<sb_idx>(pnnz, 1));
  long long permoff = 0;
  SB_CUDA(cudaMemcpyAsync(h_betajc.data(), d_betajc, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(h_ordered.data(), d_ordered, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(h_pivperm.data(), d_pivperm, sizeof(int) * h_pivperm.size(), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(h_beta.data(), d_beta, sizeof(double) * h_beta.size(), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(p_out, d_p, sizeof(double) * pnnz, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(dd.data(), d_d, sizeof(double) * dd.size(), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(&permoff, d_permoff, sizeof(long long), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  for (sb_idx i = 0; i < dznnz; i++) d_out[dzir[i]] = dd[i];            // lab(dz.ir) = d
  for (sb_idx k = 0; k <= n; k++) betajc_out[k] = h_betajc[k];
  *nbeta = h_betajc[n];
  for (sb_idx i = 0; i < *nbeta; i++) beta_out[i] = h_beta[i];
  sb_idx permnnz = 0;
  for (sb_idx k = 0; k < n; k++) { dopiv_out[k] = h_ordered[k]; permnnz += h_ordered[k] ? dzjc[k + 1] : 0; }
  // the reference advances its perm cursor only for re-ordered columns with smult != 0 (dpr1fact.c:588-596)
  for (sb_idx i = 0; i < permnnz && i < (sb_idx)h_pivperm.size(); i++) pivperm_out[i] = h_pivperm[i];
  *npivperm = permnnz;
  (void)permoff;
  return 0;
}

// y = fwdpr1(Lden,b) / bwdpr1(Lden,b): y is m x nrhs, updated in place on the rows listed in dz.ir.
int sb200_dpr1solve(int backward, sb_idx m, sb_idx nrhs, sb_idx nden, const sb_idx *dzjc, const sb_idx *dzir, const double *p,
                    const sb_idx *pivperm, sb_idx permnnz, const double *beta, const sb_idx *betajc, const double *dopiv, double *y) {
  SB_TRY(ensure_init());
  if (nden == 0 || nrhs == 0 || m == 0) return 0;
  const sb_idx dznnz = dzjc[nden];
  SB_CHECK(dznnz <= m, "Lden.dz size mismatch.");
  std::vector<int> xs(nden), ord(nden), bj(nden + 1), ir32((size_t)std::max<sb_idx>(dznnz, 1)), pp((size_t)std::max<sb_idx>(permnnz, 1));
  std::vector<long long> poff(nden + 1), permoff(nden + 1);
  sb_idx pnnz = 0, pc = 0;
  for (sb_idx k = 0; k < nden; k++) {
    xs[k] = (int)dzjc[k + 1]; poff[k] = pnnz; pnnz += dzjc[k + 1];
    ord[k] = dopiv[k] != 0.0; permoff[k] = pc; if (ord[k]) pc += dzjc[k + 1];
    bj[k] = (int)betajc[k];
  }
  bj[nden] = (int)betajc[nden];
  SB_CHECK(pc <= permnnz || pc == 0, "Lden.pivperm too short");
  for (sb_idx i = 0; i < dznnz; i++) ir32[i] = (int)dzir[i];
  for (sb_idx i = 0; i < permnnz; i++) pp[i] = (int)pivperm[i];
  arena_reset();
  cudaStream_t st = ctx().stream;
  int *d_xs = arena<int>(nden), *d_ord = arena<int>(nden), *d_bj = arena<int>(nden + 1), *d_ir = arena<int>(ir32.size()), *d_pp = arena<int>(pp.size());
  long long *d_poff = arena<long long>(nden + 1), *d_permoff = arena<long long>(nden + 1);
  double *d_p = arena<double>(std::max<sb_idx>(pnnz, 1)), *d_beta = arena<double>(std::max<sb_idx>(bj[nden], 1)), *d_y = arena<double>(m * nrhs),
         *d_f = arena<double>(std::max<sb_idx>(dznnz, 1) * nrhs);
  SB_CHECK(d_xs && d_ord && d_bj && d_ir && d_pp && d_poff && d_permoff && d_p && d_beta && d_y && d_f, "dpr1 solve: out of device memory");
  SB_CUDA(cudaMemcpyAsync(d_xs, xs.data(), sizeof(int) * nden, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_ord, ord.data(), sizeof(int) * nden, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_bj, bj.data(), sizeof(int) * (nden + 1), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_ir, ir32.data(), sizeof(int) * ir32.size(), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_pp, pp.data(), sizeof(int) * pp.size(), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_poff, poff.data(), sizeof(long long) * (nden + 1), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_permoff, permoff.data(), sizeof(long long) * (nden + 1), cudaMemcpyHostToDevice, st));
  if (pnnz) SB_CUDA(cudaMemcpyAsync(d_p, p, sizeof(double) * pnnz, cudaMemcpyHostToDevice, st));
  if (bj[nden]) SB_CUDA(cudaMemcpyAsync(d_beta, beta, sizeof(double) * bj[nden], cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_y, y, sizeof(double) * m * nrhs, cudaMemcpyHostToDevice, st));
  prodform_solve_kernel<<<(unsigned)((nrhs + 63) / 64), 64, 0, st>>>(backward, (int)nrhs, (int)m, (int)nden, (int)dznnz, d_ir, d_xs, d_poff,
                                                                      d_permoff, d_p, d_pp, d_beta, d_bj, d_ord, d_y, d_f);
  SB_LAUNCH_CHECK_N("prodform_solve_kernel");
  SB_CUDA(cudaMemcpyAsync(y, d_y, sizeof(double) * m * nrhs, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
