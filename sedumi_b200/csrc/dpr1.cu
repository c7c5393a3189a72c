// dpr1.cu -- dense-column handling: product-form LDL' of diag(d) + sum_k smult_k p_k p_k'
// (dpr1fact) and the product-form solves (fwdpr1, bwdpr1).
//
// Reference semantics:
//   dpr1fact.c:97-135    first pass over a column: pivot j is accepted while the implied multipliers
//                        stay below maxu, otherwise postponed
//   dpr1fact.c:224-240   second pass over the postponed rows, sorted by decreasing p^2
//   dpr1fact.c:280-477   dodpr1fact: dependent rows (d = 0), removal of one dependency when t > 0
//   dpr1fact.c:495-512   findnewdep after a subtraction (smult < 0, Lorentz trace columns)
//   dpr1fact.c:549-621   prodformfact: columns in order, each followed by a forward solve of the
//                        later columns that overlap it (auxfwdpr1.c:44-122)
//   fwdpr1.c:70-90, bwdpr1.c:65-162   product-form forward / backward solves
//
// GPU formulation.  The reference's loops are first-order recurrences, and first-order recurrences are scans:
//   * factor:  t_{j+1} = t_j + x_j / d_j over the ACCEPTED pivots (x = p.^2): an exclusive prefix sum gives every
//     t_j at once, from which fi_j = x_j + t_j d_j, the new d_j = fi_j / t_j and beta_j = p_j / fi_j follow
//     elementwise.  Which pivots are accepted depends on t, so the stability test is evaluated speculatively for all
//     rows under the current postponed set; the FIRST failing row is final (it only depends on earlier rows), it
//     joins the postponed set and the scan is repeated -- (number of postponed rows + 1) block-wide passes, 1 pass
//     in the common case.  Suffix maxima (mu), the running maximum over postponed rows, the stream compaction of
//     the accepted pivots and the partition around dependent rows are scans as well; the postponed rows are ranked
//     by a parallel counting sort and finished with one more prefix sum.
//   * solves:  t_{i+1} = (1 - beta_i p_i) t_i + beta_i y_i  -- an affine map per row, composed by a block-wide
//     scan of (a, b) pairs; the backward solve is the same recurrence run from the other end.
// One CTA of 1024 threads per column (factor) / per (right-hand side) or per later column (solves).  Only the
// bookkeeping of the short list of dependent rows (dpr1fact.c:362-376,495-512) is done by a single thread.
#include <algorithm>
#include "sb_internal.h"

namespace sb {

struct KD { double r; int k; };
static const int DT = 1024;

struct Aff { double a, b; };                         // t -> a t + b
__device__ __forceinline__ double shfl_up_t(double v, int o) { return __shfl_up_sync(0xffffffffu, v, o); }
__device__ __forceinline__ int shfl_up_t(int v, int o) { return __shfl_up_sync(0xffffffffu, v, o); }
__device__ __forceinline__ Aff shfl_up_t(Aff v, int o) { return Aff{__shfl_up_sync(0xffffffffu, v.a, o), __shfl_up_sync(0xffffffffu, v.b, o)}; }
struct OpSum { __device__ double operator()(double e, double l) const { return e + l; } };
struct OpMax { __device__ double operator()(double e, double l) const { return fmax(e, l); } };
struct OpAddI { __device__ int operator()(int e, int l) const { return e + l; } };
struct OpAff { __device__ Aff operator()(Aff e, Aff l) const { return Aff{l.a * e.a, l.a * e.b + l.b}; } };   // earlier, then later

// Block-wide EXCLUSIVE scan of one value per thread, in thread order; *total (optional) = reduction over the block.
template <typename T, typename Op>
__device__ __forceinline__ T block_exscan(T v, Op op, T ident, T *sh, T *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  T inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { T u = shfl_up_t(inc, o); if (lane >= o) inc = op(u, inc); }
  __syncthreads();
  if (lane == 31) sh[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    T w = lane < nw ? sh[lane] : ident;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { T u = shfl_up_t(w, o); if (lane >= o) w = op(u, w); }
    sh[lane] = w;
  }
  __syncthreads();
  T ex = shfl_up_t(inc, 1);
  if (lane == 0) ex = ident;
  const T pre = warp > 0 ? op(sh[warp - 1], ex) : ex;
  if (total) *total = sh[nw - 1];
  __syncthreads();
  return pre;
}
__device__ __forceinline__ int block_min_int(int v, int *sh) {
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_down_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_down_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) sh[0] = v;
  }
  __syncthreads();
  v = sh[0];
  __syncthreads();
  return v;
}
#define DPR1_CHUNK(n, lo, hi) const int _c = ((n) + (int)blockDim.x - 1) / (int)blockDim.x; \
  const int lo = min((n), (int)threadIdx.x * _c), hi = min((n), lo + _c)

// dependent-row bookkeeping (a short sorted list with a tail sentinel; dpr1fact.c:362-376 and :495-512)
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

struct Dpr1Work {         // scratch of one factorisation, each array at least max_k xs[k] long
  double *x, *mu, *tq;    // p.^2 ; suffix maxima in candidate order ; t at each candidate position
  int *ord, *post, *slot; // candidate order ; postponed flags ; output slot of each candidate
  KD *kd;                 // postponed rows (row, p^2)
  int *cursor;            // [0] = next free entry of beta, [1] = next free entry of pivperm, [2] = ndep, [3] = maxndep
  int *permoffs;          // per column: where its pivot order starts in pivperm (valid when ordered[k])
};

// One dense column k: factor diag(d) + smult p p' in product form (dodpr1fact, dpr1fact.c:280-477).
__global__ void __launch_bounds__(DT)
dpr1_column_kernel(int k, const int *xs, const long long *poff, const double *p, int *pivperm, double *beta, int *betajc,
                   double *d, int *ordered, const int *colperm, const double *smult, int *dep, double maxu, Dpr1Work W) {
  __shared__ double shd[33];
  __shared__ int shi[33];
  __shared__ int s_idep, s_deldep, s_case;
  __shared__ double s_h, s_t0;
  const int tid = threadIdx.x;
  const int mk = xs[k];
  const double *pk = p + poff[k];
  const double sm = smult[colperm[k]];
  const int inz = W.cursor[0];
  double *betak = beta + inz;
  int *permk = pivperm + W.cursor[1];
  if (tid == 0) { betajc[k] = inz; W.permoffs[k] = W.cursor[1]; }
  if (sm == 0.0) {                                   // diag(d) + 0 p p' = I diag(d) I  (dpr1fact.c:291-294)
    if (tid == 0) { ordered[k] = 0; betajc[k + 1] = inz; }
    return;
  }
  for (int i = tid; i < mk; i += blockDim.x) { W.x[i] = pk[i] * pk[i]; W.post[i] = 0; }
  // ---- dependent rows inside this column
  if (tid == 0) {
    int ndep = W.cursor[2];
    double t0 = 1.0 / sm;
    s_t0 = t0; s_idep = -1; s_deldep = 0; s_h = 0.0; s_case = 0;
    if (dep[0] < mk) {                               // case B (dpr1fact.c:352-410); the list is short
      s_case = 1;
      double psqrdep = 0.0; int j = 0;
      for (int i = 0; dep[i] < mk; i++) { const double v = pk[dep[i]] * pk[dep[i]]; if (v > psqrdep) { j = i; psqrdep = v; } }
      if (psqrdep > 0.0) {
        const int idep = dep[j];
        s_idep = idep;
        if (t0 > 0.0) {
          s_deldep = 1;
          for (int a = j; a < ndep; a++) dep[a] = dep[a + 1];
          s_h = maxu * maxu * psqrdep;
          dep[ndep] = idep;
          W.cursor[2] = --ndep;
        } else s_h = psqrdep;
      } else s_idep = dep[0];
    }
  }
  __syncthreads();
  const double t0 = s_t0, h = s_h;
  const int idep = s_idep, deldep = s_deldep, caseB = s_case;
  // ---- candidate order: case A = 0..mk-1; case B = [rows with x > h (ascending), idep, the rest (descending)]
  int n;
  if (!caseB) {
    for (int i = tid; i < mk; i += blockDim.x) W.ord[i] = i;
    n = mk;
    __syncthreads();
  } else {
    DPR1_CHUNK(mk, lo, hi);
    int cf = 0, cr = 0;
    for (int i = lo; i < hi; i++) { if (i == idep) continue; if (W.x[i] > h) cf++; else cr++; }
    int totf = 0;
    const int pf = block_exscan(cf, OpAddI(), 0, shi, &totf);
    const int pr = block_exscan(cr, OpAddI(), 0, shi, (int *)nullptr);
    int a = pf, r = pr;
    for (int i = lo; i < hi; i++) {
      if (i == idep) continue;
      if (W.x[i] > h) { W.ord[a] = i; permk[a] = i; a++; } else { permk[mk - 1 - r] = i; r++; }
    }
    n = totf;
    if (tid == 0) permk[n] = idep;
    __syncthreads();
  }
  // ---- mu[i] = max(h, max x[ord[i+1 .. n-1]])   (dpr1fact.c:316-319,416-419): reverse exclusive max-scan
  {
    DPR1_CHUNK(n, lo, hi);
    double mx = 0.0;
    for (int i = lo; i < hi; i++) mx = fmax(mx, W.x[W.ord[n - 1 - i]]);
    double run = fmax(h, block_exscan(mx, OpMax(), 0.0, shd, (double *)nullptr));
    for (int i = lo; i < hi; i++) { const int pos = n - 1 - i; W.mu[pos] = run; run = fmax(run, W.x[W.ord[pos]]); }
    __syncthreads();
  }
  // ---- first pass with speculation: under the current postponed set every t is a prefix sum; the first row that
  // fails the stability test is final and joins the set (dpr1fact.c:97-135,168-202)
  {
    DPR1_CHUNK(n, lo, hi);
    for (;;) {
      double sacc = 0.0, mpost = 0.0;
      for (int i = lo; i < hi; i++) { const int r = W.ord[i]; if (W.post[i]) mpost = fmax(mpost, W.x[r]); else sacc += W.x[r] / d[r]; }
      const double tpre = block_exscan(sacc, OpSum(), 0.0, shd, (double *)nullptr);
      const double mpre = block_exscan(mpost, OpMax(), 0.0, shd, (double *)nullptr);
      double t = t0 + tpre, muph2 = mpre;
      int fail = 0x7fffffff;
      for (int i = lo; i < hi; i++) {
        const int r = W.ord[i];
        if (W.post[i]) { muph2 = fmax(muph2, W.x[r]); continue; }
        const double dj = d[r], xr = W.x[r], fij = xr + t * dj, lim = maxu * fij;
        if (!(xr * fmax(muph2, W.mu[i]) <= lim * lim)) { fail = i; break; }
        W.tq[i] = t;
        t = fij / dj;
      }
      const int f = block_min_int(fail, shi);
      if (f == 0x7fffffff) break;
      if (tid == 0) W.post[f] = 1;
      __syncthreads();
    }
  }
  // ---- accepted pivots: slots by stream compaction; d, beta; postponed rows collected in order
  int nacc = 0, nph2 = 0;
  double tend;
  {
    DPR1_CHUNK(n, lo, hi);
    int ca = 0, cp = 0; double sacc = 0.0;
    for (int i = lo; i < hi; i++) { if (W.post[i]) cp++; else { ca++; const int r = W.ord[i]; sacc += W.x[r] / d[r]; } }
    double stot = 0.0;
    const int pa = block_exscan(ca, OpAddI(), 0, shi, &nacc);
    const int pp = block_exscan(cp, OpAddI(), 0, shi, &nph2);
    block_exscan(sacc, OpSum(), 0.0, shd, &stot);
    tend = t0 + stot;
    int a = pa, q = pp;
    for (int i = lo; i < hi; i++) {
      const int r = W.ord[i];
      if (W.post[i]) { W.kd[q].k = r; W.kd[q].r = W.x[r]; q++; }
      else { W.slot[i] = a; a++; }
    }
    __syncthreads();
    for (int i = lo; i < hi; i++) {
      if (W.post[i]) continue;
      const int r = W.ord[i];
      const double t = W.tq[i], fij = W.x[r] + t * d[r];
      d[r] = fij / t;                                 // d_new = d + p^2 / t
      betak[W.slot[i]] = pk[r] / fij;
      if (caseB || nph2) permk[W.slot[i]] = r;
    }
    __syncthreads();
  }
  // ---- second pass: postponed rows by decreasing p^2 (ties: smaller row first), one more prefix sum (dpr1fact.c:224-240)
  if (nph2) {
    for (int e = tid; e < nph2; e += blockDim.x) {
      const double xe = W.kd[e].r; const int re = W.kd[e].k;
      int rank = 0;
      for (int o = 0; o < nph2; o++) { const double xo = W.kd[o].r; rank += (xo > xe) || (xo == xe && W.kd[o].k < re); }
      W.slot[e] = rank;                               // slot[] is free again: accepted slots were consumed above
    }
    __syncthreads();
    for (int e = tid; e < nph2; e += blockDim.x) W.ord[W.slot[e]] = W.kd[e].k;      // sorted rows
    __syncthreads();
    DPR1_CHUNK(nph2, lo, hi);
    double sacc = 0.0;
    for (int i = lo; i < hi; i++) { const int r = W.ord[i]; sacc += W.x[r] / d[r]; }
    double stot = 0.0;
    const double tpre = block_exscan(sacc, OpSum(), 0.0, shd, &stot);
    double t = tend + tpre;
    for (int i = lo; i < hi; i++) {
      const int r = W.ord[i];
      const double dj = d[r], fij = W.x[r] + t * dj;
      d[r] = fij / t;
      betak[nacc + i] = pk[r] / fij;
      permk[nacc + i] = r;
      t = fij / dj;
    }
    tend += stot;
    __syncthreads();
  }
  if (tid == 0) {
    int nk = caseB ? n + deldep : mk;
    if (caseB && deldep) {                            // finish by pivoting on idep (dpr1fact.c:468-475)
      d[idep] = W.x[idep] / tend;
      betak[nacc + nph2] = 1.0 / pk[idep];
    }
    const int useperm = caseB || nph2 > 0;
    ordered[k] = useperm;
    W.cursor[0] = inz + nk;
    betajc[k + 1] = inz + nk;
    if (useperm) W.cursor[1] += mk;
    if (sm < 0.0) W.cursor[2] += d_findnewdep(dep, W.cursor[2], W.cursor[3], d);
  }
}

// L(p_k, beta_k) yNEW = yOLD for one vector y (auxfwdpr1.c:44-122) as a block-wide scan of affine maps.
// rows: pivot order (perm) or nullptr = natural; beta has nk entries (beta_i = 0 beyond); all mk rows are updated.
__device__ __forceinline__ void d_fw_scan(double *y, const int *perm, const double *p, const double *beta, int mk, int nk, Aff *sh) {
  if (nk < 1) return;
  DPR1_CHUNK(mk, lo, hi);
  Aff loc{1.0, 0.0};
  for (int i = lo; i < hi; i++) {
    const int r = perm ? perm[i] : i;
    const double b = i < nk ? beta[i] : 0.0;
    loc = OpAff()(loc, Aff{1.0 - b * p[r], b * y[r]});
  }
  const Aff pre = block_exscan(loc, OpAff(), Aff{1.0, 0.0}, sh, (Aff *)nullptr);
  double t = pre.b;                                    // t starts at 0
  for (int i = lo; i < hi; i++) {
    const int r = perm ? perm[i] : i;
    const double yi = y[r] - t * p[r];
    y[r] = yi;
    if (i < nk) t += yi * beta[i];
  }
  __syncthreads();
}
// L(p_k, beta_k)' yNEW = yOLD (bwdpr1.c:65-162): t starts as p(n:m-1)'y and runs from row n-1 down to 0.
__device__ __forceinline__ void d_bw_scan(double *y, const int *perm, const double *p, const double *beta, int mk, int nk, Aff *sh, double *shd) {
  if (nk < 1) return;
  double tail = 0.0;
  for (int i = nk + threadIdx.x; i < mk; i += blockDim.x) { const int r = perm ? perm[i] : i; tail += p[r] * y[r]; }
  double t0 = 0.0;
  block_exscan(tail, OpSum(), 0.0, shd, &t0);
  DPR1_CHUNK(nk, lo, hi);                              // positions in REVERSE order: q = nk-1-i
  Aff loc{1.0, 0.0};
  for (int q = lo; q < hi; q++) {
    const int i = nk - 1 - q, r = perm ? perm[i] : i;
    loc = OpAff()(loc, Aff{1.0 - p[r] * beta[i], p[r] * y[r]});
  }
  const Aff pre = block_exscan(loc, OpAff(), Aff{1.0, 0.0}, sh, (Aff *)nullptr);
  double t = pre.a * t0 + pre.b;
  for (int q = lo; q < hi; q++) {
    const int i = nk - 1 - q, r = perm ? perm[i] : i;
    const double yi = y[r] - t * beta[i];
    y[r] = yi;
    t += p[r] * yi;
  }
  __syncthreads();
}

// Forward solve of the later columns j > k that overlap column k (dpr1fact.c:578-596): one CTA per column j.
__global__ void __launch_bounds__(DT)
dpr1_fwcols_kernel(int k, int n, const int *xs, const long long *poff, double *p, const int *pivperm, const double *beta,
                   const int *betajc, const int *ordered, const int *colperm, const int *firstpiv, const double *smult,
                   const int *permoffs) {
  __shared__ Aff sha[33];
  const int j = k + 1 + blockIdx.x;
  if (j >= n || smult[colperm[k]] == 0.0 || firstpiv[colperm[j]] > k) return;
  const int nk = betajc[k + 1] - betajc[k];
  d_fw_scan(p + poff[j], ordered[k] ? pivperm + permoffs[k] : nullptr, p + poff[k], beta + betajc[k], xs[k], nk, sha);
}

// Product-form solves over all dense columns, one CTA per right-hand side (fwdpr1.c:70-90, bwdpr1.c:131-162).
__global__ void __launch_bounds__(DT)
dpr1_solve_kernel(int backward, int m, int nden, int dznnz, const int *dzir, const int *xs, const long long *poff,
                  const long long *permoff, const double *p, const int *pivperm, const double *beta, const int *betajc,
                  const int *ordered, double *y, double *fwork) {
  __shared__ Aff sha[33];
  __shared__ double shd[33];
  double *yy = y + (long long)blockIdx.x * m, *f = fwork + (long long)blockIdx.x * dznnz;
  for (int i = threadIdx.x; i < dznnz; i += blockDim.x) f[i] = yy[dzir[i]];
  __syncthreads();
  if (!backward) {
    for (int k = 0; k < nden; k++)
      d_fw_scan(f, ordered[k] ? pivperm + permoff[k] : nullptr, p + poff[k], beta + betajc[k], xs[k], betajc[k + 1] - betajc[k], sha);
  } else {
    for (int k = nden - 1; k >= 0; k--)
      d_bw_scan(f, ordered[k] ? pivperm + permoff[k] : nullptr, p + poff[k], beta + betajc[k], xs[k], betajc[k + 1] - betajc[k], sha, shd);
  }
  for (int i = threadIdx.x; i < dznnz; i += blockDim.x) yy[dzir[i]] = f[i];
}

__global__ void __launch_bounds__(DT)
dpr1_solve_dev_kernel(int backward, int m, int nden, int dznnz, const int *dzir, const int *xs, const long long *poff,
                      const int *permoffs, const double *p, const int *pivperm, const double *beta, const int *betajc,
                      const int *ordered, double *y, double *fwork) {
  __shared__ Aff sha[33];
  __shared__ double shd[33];
  double *yy = y + (long long)blockIdx.x * m, *f = fwork + (long long)blockIdx.x * dznnz;
  for (int i = threadIdx.x; i < dznnz; i += blockDim.x) f[i] = yy[dzir[i]];
  __syncthreads();
  if (!backward) {
    for (int k = 0; k < nden; k++)
      d_fw_scan(f, ordered[k] ? pivperm + permoffs[k] : nullptr, p + poff[k], beta + betajc[k], xs[k], betajc[k + 1] - betajc[k], sha);
  } else {
    for (int k = nden - 1; k >= 0; k--)
      d_bw_scan(f, ordered[k] ? pivperm + permoffs[k] : nullptr, p + poff[k], beta + betajc[k], xs[k], betajc[k + 1] - betajc[k], sha, shd);
  }
  for (int i = threadIdx.x; i < dznnz; i += blockDim.x) yy[dzir[i]] = f[i];
}

__global__ void gather_idx_kernel(int n, const int *idx, const double *src, double *dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
__global__ void dscale_kernel(int m, long long tot, const double *d, const int *flag, const double *lb, double *y) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= tot) return;
  const int k = (int)(i % m);
  double dk = d[k];
  if (flag && lb && flag[k] == 1 && dk <= lb[k]) dk = 1.0;
  y[i] /= dk;
}

// ---- device-resident chain: p from the dense block of L\Ad, d in the dz order, the list of dependent rows
__global__ void dpr1_gather_kernel(int m, int n, const int *xs, const long long *poff, const int *colperm, const int *dzir,
                                   const double *LAD, double *p) {
  const int k = blockIdx.y;
  const double *col = LAD + (long long)colperm[k] * m;
  double *pk = p + poff[k];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < xs[k]; i += gridDim.x * blockDim.x) pk[i] = col[dzir[i]];
}
// dd = d_in(dz.ir); dep = ascending positions with dd <= 0, tail sentinel m; cursor = {0, 0, ndep, ndep}
__global__ void __launch_bounds__(DT) dpr1_prepare_kernel(int m, int dznnz, const int *dzir, const double *d_in, double *dd, int *dep, int *cursor) {
  __shared__ int shi[33];
  DPR1_CHUNK(dznnz, lo, hi);
  int c = 0;
  for (int i = lo; i < hi; i++) { const double v = d_in[dzir[i]]; dd[i] = v; c += (v <= 0.0); }
  int ndep = 0;
  int pos = block_exscan(c, OpAddI(), 0, shi, &ndep);
  for (int i = lo; i < hi; i++) if (dd[i] <= 0.0) dep[pos++] = i;
  if (threadIdx.x == 0) { dep[ndep] = m; cursor[0] = 0; cursor[1] = 0; cursor[2] = ndep; cursor[3] = ndep; }
}
__global__ void dpr1_scatter_d_kernel(int m, int dznnz, const int *dzir, const double *d_in, const double *dd, double *d_out, int phase) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (phase == 0) { if (i < m) d_out[i] = d_in[i]; }
  else if (i < dznnz) d_out[dzir[i]] = dd[i];
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
      *d_betajc = arena<int>(n + 1), *d_ordered = arena<int>(n), *d_dep = arena<int>(m + 2), *d_cursor = arena<int>(4), *d_permoffs = arena<int>(n);
  long long *d_poff = arena<long long>(n + 1);
  const size_t wl = (size_t)std::max<sb_idx>(dznnz, 1);
  double *d_p = arena<double>(std::max<sb_idx>(pnnz, 1)), *d_beta = arena<double>(std::max<sb_idx>(pnnz, 1)), *d_d = arena<double>(wl),
         *d_smult = arena<double>(n);
  Dpr1Work W;
  W.x = arena<double>(wl); W.mu = arena<double>(wl); W.tq = arena<double>(wl);
  W.ord = arena<int>(wl); W.post = arena<int>(wl); W.slot = arena<int>(wl); W.kd = arena<KD>(wl);
  W.cursor = d_cursor; W.permoffs = d_permoffs;
  SB_CHECK(d_xs && d_colp && d_first && d_pivperm && d_betajc && d_ordered && d_dep && d_cursor && d_permoffs && d_poff && d_p && d_beta &&
           d_d && d_smult && W.x && W.mu && W.tq && W.ord && W.post && W.slot && W.kd, "dpr1fact: out of device memory");
  const int cursor0[4] = {0, 0, ndep, ndep};
  SB_CUDA(cudaMemcpyAsync(d_xs, xs.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_colp, colp.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_first, firstp.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_poff, poff.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_p, p.data(), sizeof(double) * p.size(), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_d, dd.data(), sizeof(double) * dd.size(), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_smult, smult, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_dep, dep.data(), sizeof(int) * (m + 2), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(d_cursor, cursor0, sizeof(cursor0), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemsetAsync(d_beta, 0, sizeof(double) * std::max<sb_idx>(pnnz, 1), st));
  SB_CUDA(cudaMemsetAsync(d_betajc, 0, sizeof(int) * (n + 1), st));
  for (sb_idx k = 0; k < n; k++) {
    dpr1_column_kernel<<<1, DT, 0, st>>>((int)k, d_xs, d_poff, d_p, d_pivperm, d_beta, d_betajc, d_d, d_ordered, d_colp, d_smult, d_dep, maxu, W);
    SB_LAUNCH_CHECK_N("dpr1_column_kernel");
    if (k + 1 < n) {
      dpr1_fwcols_kernel<<<(unsigned)(n - k - 1), DT, 0, st>>>((int)k, (int)n, d_xs, d_poff, d_p, d_pivperm, d_beta, d_betajc, d_ordered, d_colp,
                                                               d_first, d_smult, d_permoffs);
      SB_LAUNCH_CHECK_N("dpr1_fwcols_kernel");
    }
  }
  std::vector<int> h_betajc(n + 1), h_ordered(n), h_pivperm((size_t)std::max<sb_idx>(pnnz, 1));
  std::vector<double> h_beta((size_t)std::max<sb_idx>(pnnz, 1));
  SB_CUDA(cudaMemcpyAsync(h_betajc.data(), d_betajc, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(h_ordered.data(), d_ordered, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(h_pivperm.data(), d_pivperm, sizeof(int) * h_pivperm.size(), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(h_beta.data(), d_beta, sizeof(double) * h_beta.size(), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(p_out, d_p, sizeof(double) * pnnz, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(dd.data(), d_d, sizeof(double) * dd.size(), cudaMemcpyDeviceToHost, st));
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
  dpr1_solve_kernel<<<(unsigned)nrhs, DT, 0, st>>>(backward, (int)m, (int)nden, (int)dznnz, d_ir, d_xs, d_poff, d_permoff, d_p, d_pp, d_beta,
                                                    d_bj, d_ord, d_y, d_f);
  SB_LAUNCH_CHECK_N("dpr1_solve_kernel");
  SB_CUDA(cudaMemcpyAsync(y, d_y, sizeof(double) * m * nrhs, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------ device-resident product form (HotPath chain)
struct sb200_dpr1_plan {
  int m = 0, n = 0, dznnz = 0;
  long long pnnz = 0;
  int solve_nrhs = 0;
  sb::DevBuf<int> d_xs, d_colp, d_first, d_dzir, d_pivperm, d_betajc, d_ordered, d_dep, d_cursor, d_permoffs, d_ord, d_post, d_slot;
  sb::DevBuf<long long> d_poff, d_permoff64;
  sb::DevBuf<double> d_p, d_beta, d_d, d_x, d_mu, d_tq, d_f;
  sb::DevBuf<sb::KD> d_kd;
};

extern "C" {

// Lsymb of the dense columns (symbcholden.m:45-62): dz (m x n pattern, jc/ir), perm (column order), first -- all 0-based.
int sb200_dpr1_plan_create(sb200_dpr1_plan **plan, sb_idx m, sb_idx n, const sb_idx *dzjc, const sb_idx *dzir, const sb_idx *colperm,
                           const sb_idx *firstpiv) {
  SB_TRY(ensure_init());
  sb200_dpr1_plan *pl = new sb200_dpr1_plan();
  pl->m = (int)m; pl->n = (int)n; pl->dznnz = (int)dzjc[n];
  std::vector<int> xs(n), colp(n), firstp(n), ir32((size_t)std::max<sb_idx>(dzjc[n], 1));
  std::vector<long long> poff(n + 1, 0);
  long long pnnz = 0;
  for (sb_idx k = 0; k < n; k++) { xs[k] = (int)dzjc[k + 1]; poff[k] = pnnz; pnnz += dzjc[k + 1]; colp[k] = (int)colperm[k]; firstp[k] = (int)firstpiv[k]; }
  poff[n] = pnnz; pl->pnnz = pnnz;
  for (sb_idx i = 0; i < dzjc[n]; i++) ir32[i] = (int)dzir[i];
  const size_t wl = (size_t)std::max(pl->dznnz, 1), pl1 = (size_t)std::max<long long>(pnnz, 1);
  int rc = pl->d_xs.upload(xs) || pl->d_colp.upload(colp) || pl->d_first.upload(firstp) || pl->d_dzir.upload(ir32) || pl->d_poff.upload(poff) ||
           pl->d_pivperm.alloc(pl1) || pl->d_betajc.alloc(n + 1) || pl->d_ordered.alloc(n) || pl->d_dep.alloc(m + 2) || pl->d_cursor.alloc(4) ||
           pl->d_permoffs.alloc(n) || pl->d_permoff64.alloc(n + 1) || pl->d_ord.alloc(wl) || pl->d_post.alloc(wl) || pl->d_slot.alloc(wl) || pl->d_p.alloc(pl1) ||
           pl->d_beta.alloc(pl1) || pl->d_d.alloc(wl) || pl->d_x.alloc(wl) || pl->d_mu.alloc(wl) || pl->d_tq.alloc(wl) || pl->d_kd.alloc(wl);
  if (rc) { delete pl; return 1; }
  SB_CUDA(cudaStreamSynchronize(ctx().stream));
  *plan = pl;
  return 0;
}
void sb200_dpr1_plan_destroy(sb200_dpr1_plan *pl) { delete pl; }

// [Lden, d] = dpr1fact(L\Ad, L.d, Lsymb, smult, maxu) on device data (deninfac.m:67-72).  LAD_dev: m x n, column-major,
// rows in the factor's permuted order (what sb200_fwblkslv_dev returns for the dense columns); d_in/d_out: m doubles.
int sb200_dpr1fact_dev(sb200_dpr1_plan *pl, const double *LAD_dev, const double *smult_dev, double maxu, const double *d_in_dev,
                       double *d_out_dev) {
  SB_TRY(ensure_init());
  cudaStream_t st = ctx().stream;
  const int n = pl->n, m = pl->m;
  if (n == 0) return 0;
  int maxx = pl->dznnz;
  dpr1_gather_kernel<<<dim3((unsigned)std::max(1, std::min((maxx + 255) / 256, 64)), (unsigned)n), 256, 0, st>>>(m, n, pl->d_xs.p, pl->d_poff.p, pl->d_colp.p,
                                                                                                     pl->d_dzir.p, LAD_dev, pl->d_p.p);
  SB_LAUNCH_CHECK_N("dpr1_gather_kernel");
  dpr1_prepare_kernel<<<1, DT, 0, st>>>(m, pl->dznnz, pl->d_dzir.p, d_in_dev, pl->d_d.p, pl->d_dep.p, pl->d_cursor.p);
  SB_LAUNCH_CHECK_N("dpr1_prepare_kernel");
  SB_CUDA(cudaMemsetAsync(pl->d_beta.p, 0, sizeof(double) * std::max<long long>(pl->pnnz, 1), st));
  Dpr1Work W;
  W.x = pl->d_x.p; W.mu = pl->d_mu.p; W.tq = pl->d_tq.p; W.ord = pl->d_ord.p; W.post = pl->d_post.p; W.slot = pl->d_slot.p; W.kd = pl->d_kd.p;
  W.cursor = pl->d_cursor.p; W.permoffs = pl->d_permoffs.p;
  for (int k = 0; k < n; k++) {
    dpr1_column_kernel<<<1, DT, 0, st>>>(k, pl->d_xs.p, pl->d_poff.p, pl->d_p.p, pl->d_pivperm.p, pl->d_beta.p, pl->d_betajc.p, pl->d_d.p,
                                         pl->d_ordered.p, pl->d_colp.p, smult_dev, pl->d_dep.p, maxu, W);
    SB_LAUNCH_CHECK_N("dpr1_column_kernel");
    if (k + 1 < n) {
      dpr1_fwcols_kernel<<<(unsigned)(n - k - 1), DT, 0, st>>>(k, n, pl->d_xs.p, pl->d_poff.p, pl->d_p.p, pl->d_pivperm.p, pl->d_beta.p, pl->d_betajc.p,
                                                               pl->d_ordered.p, pl->d_colp.p, pl->d_first.p, smult_dev, pl->d_permoffs.p);
      SB_LAUNCH_CHECK_N("dpr1_fwcols_kernel");
    }
  }
  dpr1_scatter_d_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(m, pl->dznnz, pl->d_dzir.p, d_in_dev, pl->d_d.p, d_out_dev, 0);
  SB_LAUNCH_CHECK_N("dpr1_scatter_d_kernel");
  dpr1_scatter_d_kernel<<<(unsigned)((std::max(pl->dznnz, 1) + 255) / 256), 256, 0, st>>>(m, pl->dznnz, pl->d_dzir.p, d_in_dev, pl->d_d.p, d_out_dev, 1);
  SB_LAUNCH_CHECK_N("dpr1_scatter_d_kernel");
  return 0;
}

// y = fwdpr1(Lden, y) / bwdpr1(Lden, y) in place on m x nrhs device data, with the factor the plan holds.
int sb200_dpr1solve_dev(sb200_dpr1_plan *pl, int backward, double *y_dev, sb_idx nrhs) {
  SB_TRY(ensure_init());
  if (pl->n == 0 || nrhs == 0) return 0;
  if (pl->solve_nrhs < nrhs) {
    SB_CHECK(!ctx().capturing, "dpr1 solve: workspace must be sized before graph capture");
    SB_TRY(pl->d_f.alloc((size_t)std::max(pl->dznnz, 1) * nrhs));
    pl->solve_nrhs = (int)nrhs;
  }
  dpr1_solve_dev_kernel<<<(unsigned)nrhs, DT, 0, ctx().stream>>>(backward, pl->m, pl->n, pl->dznnz, pl->d_dzir.p, pl->d_xs.p, pl->d_poff.p, pl->d_permoffs.p,
                                                                 pl->d_p.p, pl->d_pivperm.p, pl->d_beta.p, pl->d_betajc.p, pl->d_ordered.p, y_dev, pl->d_f.p);
  SB_LAUNCH_CHECK_N("dpr1_solve_kernel");
  return 0;
}
// smult = d.l(dense.cols) (deninfac.m:61): dst[i] = src[idx[i]]
int sb200_gather_dev(sb_idx n, const int *idx_dev, const double *src_dev, double *dst_dev) {
  SB_TRY(ensure_init());
  if (n <= 0) return 0;
  gather_idx_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx().stream>>>((int)n, idx_dev, src_dev, dst_dev);
  SB_LAUNCH_CHECK_N("gather_idx_kernel");
  return 0;
}
// y ./= d with deninfac's repair of the pivots that are still zero (deninfac.m:88-93): d <= lb on a skipped pivot -> 1
int sb200_scale_by_d_dev(sb_idx m, sb_idx nrhs, const double *d_dev, const int *flag_dev, const double *lb_dev, double *y_dev) {
  SB_TRY(ensure_init());
  const long long tot = (long long)m * nrhs;
  if (tot <= 0) return 0;
  dscale_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx().stream>>>((int)m, tot, d_dev, flag_dev, lb_dev, y_dev);
  SB_LAUNCH_CHECK_N("dscale_kernel");
  return 0;
}
// the factor as the MEX interface carries it (for tests): sizes first (p, beta, pivperm lengths), then the arrays
int sb200_dpr1_plan_download(sb200_dpr1_plan *pl, double *p, double *beta, int *betajc, int *pivperm, int *ordered) {
  cudaStream_t st = ctx().stream;
  SB_CUDA(cudaMemcpyAsync(p, pl->d_p.p, sizeof(double) * pl->pnnz, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(beta, pl->d_beta.p, sizeof(double) * pl->pnnz, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(betajc, pl->d_betajc.p, sizeof(int) * (pl->n + 1), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(pivperm, pl->d_pivperm.p, sizeof(int) * pl->pnnz, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(ordered, pl->d_ordered.p, sizeof(int) * pl->n, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
