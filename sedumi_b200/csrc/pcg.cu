// pcg.cu -- the direct step of wrapPcg on the device: the scaling operations, the two products with
// the constraint matrix, the factor solves and the residual of one search-direction computation, chained on the library
// stream without a host round trip.
//
// Reference semantics (wrapPcg.m:42-97, Amul.m:42-56, vecsym.c, psdscale.m):
//     dx  = D' rv            D' = [sqrt(d.l) .* ; psdscale(d, ., K, 1)]
//     r   = A dx + rb        (Amul: (x' At)')
//     p   = L \ r ; y = p ./ L.d ; ssqrNew = p' y ; p = L' \ y
//     x   = vecsym(At p) ;  dx2 = D x ;  ssqrdx = |dx2|^2 ;  alpha = ssqrNew / ssqrdx
//     y   = alpha p ;  dx = rv - alpha dx2
//     r   = A D' dx + rb ;  normr = |r|_inf
// Scope: LP + PSD cones without dense columns (the Lorentz terms of asmDxq.m and the dense-column products of Amul.m
// are refused, not approximated); the PCG refinement loop (loopPcg.m) stays with the caller, who sees normr.
// Scalars live in device memory (scal[0..3] = ssqrNew, ssqrdx, alpha, normr), so nothing here synchronises.
#include "sb_internal.h"

extern "C" int sb200_ada_plan_csr(sb200_ada_plan *plan, const long long **Ajc, const int **Air, const double **Apr,
                                  const long long **rowptr, const int **rowcol, const int **rowsrc, sb_idx *N, sb_idx *m,
                                  sb_idx *lpN, sb_idx *nq);
extern "C" int sb200_psd_plan_blocks(sb200_psd_plan *plan, const int **n_dev, const long long **off_dev, int *nblk, int *maxn);
extern "C" int sb200_ldl_solve2_dev(sb200_chol_plan *plan, const double *Lrect_dev, const double *d_dev, const int *flag_dev,
                                    const double *b_dev, double *w_dev, double *y_dev, sb_idx nrhs, double *ssqr_dev);

namespace sb {

// y(j) = sum_r At(r,j) x(r) (+ add(j)): one warp per column of At
__global__ void __launch_bounds__(256)
pcg_at_dot_kernel(int m, const long long *Ajc, const int *Air, const double *Apr, const double *x, const double *add, double *y) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= m) return;
  double acc = 0.0;
  for (long long p = Ajc[j] + lane; p < Ajc[j + 1]; p += 32) acc += Apr[p] * x[Air[p]];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) y[j] = acc + (add ? add[j] : 0.0);
}
// y(r) = sum_j At(r,j) p(j): row-wise gather through the CSR copy of the pattern (deterministic, no atomics)
__global__ void pcg_at_mul_kernel(long long N, const long long *rowptr, const int *rowcol, const int *rowsrc, const double *Apr,
                                  const double *p, double *y) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < N; r += (long long)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (long long t = rowptr[r]; t < rowptr[r + 1]; t++) acc += Apr[rowsrc[t]] * p[rowcol[t]];
    y[r] = acc;
  }
}
// vecsym.c:60-76 on every PSD block: Y = (X + X')/2, in place (each unordered pair is owned by its lower entry)
__global__ void pcg_vecsym_kernel(int nblk, const int *bn, const long long *boff, double *x) {
  const int k = blockIdx.y, n = bn[k];
  double *X = x + boff[k];
  const long long tot = (long long)n * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < tot; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % n), j = (int)(idx / n);
    if (i > j) { const double v = (X[idx] + X[j + (long long)i * n]) / 2; X[idx] = v; X[j + (long long)i * n] = v; }
  }
}
__global__ void pcg_lp_scale_kernel(int n, const double *dl, const double *x, double *y) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = sqrt(dl[i]) * x[i];
}
// deterministic reductions: per-block partials in a fixed partition, then one block sums them in order
template <int OP>   // 0: sum of squares, 1: max |.|
__global__ void __launch_bounds__(256) pcg_reduce1_kernel(long long n, const double *x, double *part) {
  __shared__ double sh[8];
  const long long per = (n + gridDim.x - 1) / gridDim.x, lo = blockIdx.x * per, hi = min(n, lo + per);
  double a = 0.0;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) { const double v = x[i]; a = OP == 0 ? a + v * v : fmax(a, fabs(v)); }
  for (int o = 16; o > 0; o >>= 1) { const double b = __shfl_down_sync(0xffffffffu, a, o); a = OP == 0 ? a + b : fmax(a, b); }
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) { double t = sh[0]; for (int w = 1; w < 8; w++) t = OP == 0 ? t + sh[w] : fmax(t, sh[w]); part[blockIdx.x] = t; }
}
template <int OP>
__global__ void pcg_reduce2_kernel(int np, const double *part, double *out, const double *num, double *ratio) {
  if (threadIdx.x == 0) {
    double t = part[0];
    for (int i = 1; i < np; i++) t = OP == 0 ? t + part[i] : fmax(t, part[i]);
    *out = t;
    if (ratio) *ratio = t > 0.0 ? *num / t : 0.0;          // alpha = ssqrNew / ssqrdx (0 when dx vanishes, wrapPcg.m:69-74)
  }
}
// y = alpha p ;  dx = rv - alpha dx2
__global__ void pcg_step_kernel(long long N, int m, const double *alpha, const double *p, double *y, const double *rv, const double *dx2, double *dx) {
  const double a = *alpha;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N + m; i += (long long)gridDim.x * blockDim.x) {
    if (i < m) y[i] = a * p[i];
    else { const long long t = i - m; dx[t] = rv[t] - a * dx2[t]; }
  }
}

}  // namespace sb
using namespace sb;

extern "C" {

// work_dev: 3 N + 2 m + 600 doubles of scratch.  scal_dev[0..3] = ssqrNew, ssqrdx, alpha, normr.
int sb200_wrappcg_dev(sb200_ada_plan *ap, sb200_psd_plan *pp, sb200_chol_plan *cp, const double *dl_dev, const double *u_dev,
                      const int *perm_dev, const double *Lrect_dev, const double *Ld_dev, const int *flag_dev, const double *rv_dev,
                      const double *rb_dev, double *y_dev, double *dx_dev, double *r_dev, double *scal_dev, double *work_dev) {
  SB_TRY(ensure_init());
  const long long *Ajc, *rowptr; const int *Air, *rowcol, *rowsrc; const double *Apr;
  sb_idx N, m, lpN, nq;
  SB_TRY(sb200_ada_plan_csr(ap, &Ajc, &Air, &Apr, &rowptr, &rowcol, &rowsrc, &N, &m, &lpN, &nq));
  SB_CHECK(nq == 0, "wrappcg_dev: Lorentz cones are not handled on the device (asmDxq.m)");
  const sb_idx lenud = sb200_psd_plan_lenud(pp);
  SB_CHECK(lpN + lenud == N, "wrappcg_dev: cone layout does not match At (%lld + %lld != %lld)", (long long)lpN, (long long)lenud, (long long)N);
  cudaStream_t st = ctx().stream;
  double *t1 = work_dev, *t2 = t1 + N, *t3 = t2 + N, *pv = t3 + N, *wv = pv + m, *part = wv + m;
  const int NP = 256;
  auto grid = [](long long n) { return (unsigned)std::max<long long>(1, std::min<long long>((n + 255) / 256, 2048)); };
  auto scaleD = [&](const double *x, double *y, int transp) -> int {          // y = D x resp. D' x
    if (lpN) { pcg_lp_scale_kernel<<<grid(lpN), 256, 0, st>>>((int)lpN, dl_dev, x, y); SB_LAUNCH_CHECK_N("pcg_lp_scale_kernel"); }
    if (lenud) SB_TRY(sb200_psdscale_dev(pp, u_dev, perm_dev, x + lpN, transp, y + lpN));
    return 0;
  };
  auto Adot = [&](const double *x, double *y) -> int {                        // y = A x + rb
    pcg_at_dot_kernel<<<(unsigned)((m + 7) / 8), 256, 0, st>>>((int)m, Ajc, Air, Apr, x, rb_dev, y);
    SB_LAUNCH_CHECK_N("pcg_at_dot_kernel");
    return 0;
  };
  // dx = D' rv ; r = A dx + rb
  SB_TRY(scaleD(rv_dev, t1, 1));
  SB_TRY(Adot(t1, r_dev));
  // p = L' \ ((L \ r) ./ d), ssqrNew
  SB_TRY(sb200_ldl_solve2_dev(cp, Lrect_dev, Ld_dev, flag_dev, r_dev, wv, pv, 1, scal_dev + 0));
  // x = vecsym(At p) ; dx2 = D x ; ssqrdx ; alpha
  pcg_at_mul_kernel<<<grid(N), 256, 0, st>>>(N, rowptr, rowcol, rowsrc, Apr, pv, t1);
  SB_LAUNCH_CHECK_N("pcg_at_mul_kernel");
  if (lenud) {
    const int *bn; const long long *boff; int nblk, maxn;
    SB_TRY(sb200_psd_plan_blocks(pp, &bn, &boff, &nblk, &maxn));
    dim3 g((unsigned)std::min<long long>(((long long)maxn * maxn + 255) / 256, 1024), (unsigned)nblk);
    pcg_vecsym_kernel<<<g, 256, 0, st>>>(nblk, bn, boff, t1 + lpN);
    SB_LAUNCH_CHECK_N("pcg_vecsym_kernel");
  }
  SB_TRY(scaleD(t1, t2, 0));
  pcg_reduce1_kernel<0><<<NP, 256, 0, st>>>(N, t2, part);
  SB_LAUNCH_CHECK_N("pcg_reduce_kernel");
  pcg_reduce2_kernel<0><<<1, 32, 0, st>>>(NP, part, scal_dev + 1, scal_dev + 0, scal_dev + 2);
  SB_LAUNCH_CHECK_N("pcg_reduce_kernel");
  // y = alpha p ; dx = rv - alpha dx2
  pcg_step_kernel<<<grid(N + m), 256, 0, st>>>(N, (int)m, scal_dev + 2, pv, y_dev, rv_dev, t2, dx_dev);
  SB_LAUNCH_CHECK_N("pcg_step_kernel");
  // r = A D' dx + rb ; normr
  SB_TRY(scaleD(dx_dev, t3, 1));
  SB_TRY(Adot(t3, r_dev));
  pcg_reduce1_kernel<1><<<NP, 256, 0, st>>>(m, r_dev, part + NP);
  SB_LAUNCH_CHECK_N("pcg_reduce_kernel");
  pcg_reduce2_kernel<1><<<1, 32, 0, st>>>(NP, part + NP, scal_dev + 3, nullptr, nullptr);
  SB_LAUNCH_CHECK_N("pcg_reduce_kernel");
  return 0;
}

}  // extern "C"
