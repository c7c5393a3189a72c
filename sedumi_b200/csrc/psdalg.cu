// psdalg.cu -- PSD-block algebra of the scaling update: vecsym, sqrtinv, qrK.
//
// Reference semantics:
//   vecsym.c:60-76,79-96,110-133   Y_k = (X_k + X_k')/2 on real blocks; Hermitian blocks: symmetric part of Re, skew part of Im
//   sqrtinv.c:51-78,133-148        Y_k = (Q_k / diag(sqrt(v_k)))'  (Hermitian: conjugate transpose)
//   qrK.c:86-122,263-268           per real block the Householder QR  X_k = Q_k R_k: reflector k in column k of q (rows
//                                  k..n-1), beta_k in the last column of q -- the product form psdframeit/psdinvjmul read
//                                  (vfrm.s of wregion.m); triu(r) = R_k, the part below the diagonal is "undefined" in the
//                                  reference and is left as the algorithm leaves it here too.
// vecsym and sqrtinv are streams (one thread per entry).  qrK is n-1 dependent reflections per block: one CTA per block,
// the reflector in shared memory, a block reduction for its norm, then one warp per remaining column (dot + axpy).
#include <algorithm>
#include "gemm.cuh"
#include "sb_internal.h"

namespace sb {

struct AlgBlk { int n, cplx; long long off, voff; };      // off: first entry of the block in a lenud vector; voff: in sum(n)

__global__ void vecsym_kernel(const AlgBlk *blks, const double *x, double *y) {
  const AlgBlk B = blks[blockIdx.y];
  const int n = B.n;
  const long long tot = (long long)n * n;
  const double *X = x + B.off; double *Y = y + B.off;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < tot; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % n), j = (int)(idx / n);
    const long long tr = j + (long long)i * n;
    Y[idx] = (i == j) ? X[idx] : (X[idx] + X[tr]) / 2;
    if (B.cplx) Y[tot + idx] = (i == j) ? 0.0 : (X[tot + idx] - X[tot + tr]) / 2;       // skewproj: y(i,j) = (x(i,j) - x(j,i))/2
  }
}
// y(i,j) = q(j,i) / sqrt(v_i)   [Hermitian: Im y(i,j) = -Im q(j,i) / sqrt(v_i)]
__global__ void sqrtinv_kernel(const AlgBlk *blks, const double *q, const double *v, double *y) {
  const AlgBlk B = blks[blockIdx.y];
  const int n = B.n;
  const long long tot = (long long)n * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < tot; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % n), j = (int)(idx / n);
    const double sv = sqrt(v[B.voff + i]);
    const long long src = j + (long long)i * n;
    y[B.off + idx] = q[B.off + src] / sv;
    if (B.cplx) y[B.off + tot + idx] = -q[B.off + tot + src] / sv;
  }
}

// Householder QR of one real block per CTA (qrK.c:86-122).  r: in = X, out = R in the upper triangle.
__global__ void __launch_bounds__(256) qrk_kernel(const AlgBlk *blks, double *q, double *r) {
  extern __shared__ double qk[];                    // current reflector (n doubles)
  __shared__ double sh[8], s_beta;
  const AlgBlk B = blks[blockIdx.x];
  const int n = B.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double *U = r + B.off, *Q = q + B.off;
  double *beta = Q + (long long)n * (n - 1);          // last column of q holds beta_0 .. beta_{n-2}
  for (long long idx = tid; idx < (long long)n * n; idx += blockDim.x) Q[idx] = 0.0;
  __syncthreads();
  for (int k = 0; k < n - 1; k++) {
    double *uk = U + (long long)k * n;
    double sq = 0.0;
    for (int i = k + tid; i < n; i += blockDim.x) { const double v = uk[i]; qk[i] = v; sq += v * v; }
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_down_sync(0xffffffffu, sq, o);
    if (lane == 0) sh[warp] = sq;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0; for (int w = 0; w < 8; w++) t += sh[w];
      double qkk = qk[k];
      const double dk = (qkk >= 0.0 ? 1.0 : -1.0) * sqrt(t);       // SIGN(x) = x >= 0 ? 1 : -1
      qkk += dk;
      double bk = dk * qkk;
      qk[k] = qkk;
      if (bk == 0.0) bk = 1.0;                          // all-zero column: beta = 1 (qrK.c:103-104)
      beta[k] = bk;
      uk[k] = -dk;
      s_beta = -bk;
    }
    __syncthreads();
    const double nb = s_beta;
    for (int i = k + tid; i < n; i += blockDim.x) Q[(long long)k * n + i] = qk[i];
    // reflect the columns to the right: x_i -= (q_k' x_i / beta_k) q_k
    for (int c = k + 1 + warp; c < n; c += 8) {
      double *uc = U + (long long)c * n;
      double dot = 0.0;
      for (int i = k + lane; i < n; i += 32) dot += qk[i] * uc[i];
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      const double f = dot / nb;
      for (int i = k + lane; i < n; i += 32) uc[i] += f * qk[i];
    }
    __syncthreads();
  }
}

// ---- psdjmul.m / triumtriu.m / psdfactor.m / psdinvscale.m (real blocks): transposes, symmetrisations, a per-block
// Cholesky with MATLAB's "not positive definite" flag and upper-triangular solves with many right-hand sides.
__global__ void alg_transpose_kernel(const AlgBlk *blks, const double *x, double *y) {
  __shared__ double tile[32][33];
  const AlgBlk B = blks[blockIdx.z];
  const int n = B.n, i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  if (i0 >= n || j0 >= n) return;
  const double *X = x + B.off; double *Y = y + B.off;
  for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
    const int i = i0 + threadIdx.x, j = j0 + jj;
    tile[jj][threadIdx.x] = (i < n && j < n) ? X[i + (long long)j * n] : 0.0;
  }
  __syncthreads();
  for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
    const int i = j0 + threadIdx.x, j = i0 + jj;          // Y(i,j) = X(j,i)
    if (i < n && j < n) Y[i + (long long)j * n] = tile[threadIdx.x][jj];
  }
}
// mode 0: Z = (P + P')/2 (psdjmul.m:67);  mode 1: Z = P + triu(P,1)' with P upper triangular (triumtriu.m:66);
// mode 2: Z = L + tril(L,-1)' from the lower triangle (psdfactor.m:74)
__global__ void alg_sym_kernel(const AlgBlk *blks, const double *p, double *z, int mode) {
  const AlgBlk B = blks[blockIdx.y];
  const int n = B.n;
  const long long tot = (long long)n * n;
  const double *P = p + B.off; double *Z = z + B.off;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < tot; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % n), j = (int)(idx / n);
    const long long tr = j + (long long)i * n;
    if (mode == 0) Z[idx] = 0.5 * (P[idx] + P[tr]);
    else if (mode == 1) Z[idx] = (i <= j) ? P[idx] : P[tr];
    else Z[idx] = (i >= j) ? P[idx] : P[tr];
  }
}
// chol(X,'lower') of one block per CTA, left-looking, in place on the lower triangle of W (a copy of X).
// flag[k] = 0 or the (1-based) column at which the pivot was not positive (MATLAB's second output of chol).
__global__ void __launch_bounds__(256) alg_chol_kernel(const AlgBlk *blks, double *w, int *flag) {
  extern __shared__ double rowj[];                   // L(j, 0..j-1)
  __shared__ double s_piv;
  const AlgBlk B = blks[blockIdx.x];
  const int n = B.n;
  double *L = w + B.off;
  if (threadIdx.x == 0) flag[blockIdx.x] = 0;
  for (int j = 0; j < n; j++) {
    for (int k = threadIdx.x; k < j; k += blockDim.x) rowj[k] = L[j + (long long)k * n];
    __syncthreads();
    for (int r = j + threadIdx.x; r < n; r += blockDim.x) {
      double acc = L[r + (long long)j * n];
      for (int k = 0; k < j; k++) acc -= L[r + (long long)k * n] * rowj[k];
      L[r + (long long)j * n] = acc;
      if (r == j) s_piv = acc;
    }
    __syncthreads();
    const double piv = s_piv;
    if (!(piv > 0.0)) { if (threadIdx.x == 0) flag[blockIdx.x] = j + 1; return; }
    const double sq = sqrt(piv);
    for (int r = j + threadIdx.x; r < n; r += blockDim.x) L[r + (long long)j * n] = (r == j) ? sq : L[r + (long long)j * n] / sq;
    __syncthreads();
  }
}
// Solve triu(T) Z = B for the columns of B, in place: one warp per column, the column in registers (lanes over rows),
// column k of T (contiguous) read once per elimination step.
static const int ALG_SOLVE_MAXCH = 16;               // up to 512 rows per block in registers
__global__ void __launch_bounds__(256) alg_trsolve_kernel(const AlgBlk *blks, const double *t, double *b) {
  const AlgBlk B = blks[blockIdx.y];
  const int n = B.n, lane = threadIdx.x & 31;
  const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (col >= n) return;
  const double *T = t + B.off;
  double *z = b + B.off + (long long)col * n;
  double v[ALG_SOLVE_MAXCH];
#pragma unroll
  for (int c = 0; c < ALG_SOLVE_MAXCH; c++) { const int r = lane + 32 * c; v[c] = r < n ? z[r] : 0.0; }
  for (int k = n - 1; k >= 0; k--) {
    const double *Tk = T + (long long)k * n;
    double zk = 0.0;
#pragma unroll
    for (int c = 0; c < ALG_SOLVE_MAXCH; c++) if ((k >> 5) == c) zk = v[c];
    zk = __shfl_sync(0xffffffffu, zk, k & 31) / Tk[k];
#pragma unroll
    for (int c = 0; c < ALG_SOLVE_MAXCH; c++) {
      const int r = lane + 32 * c;
      if (r == k) v[c] = zk;
      else if (r < k) v[c] -= Tk[r] * zk;
    }
  }
#pragma unroll
  for (int c = 0; c < ALG_SOLVE_MAXCH; c++) { const int r = lane + 32 * c; if (r < n) z[r] = v[c]; }
}

// Symmetric eigenvalue problem of every block: cyclic two-sided Jacobi in the round-robin (tournament) ordering, one
// CTA per block.  A step rotates m/2 disjoint index pairs at once: the rotation of pair K = (p,q) is fixed from
// (a_pp, a_qq, a_pq) at the start of the step, and every 2 x 2 sub-block A([p q],[r s]) of pairs (K, L) becomes
// J_K' A([p q],[r s]) J_L -- read and written by ONE thread, so the step is in place with two block barriers.  For odd n a
// dummy index pairs with one real index per step (identity rotation).  V accumulates the rotations column-wise.
// Stops when the off-diagonal Frobenius norm (summed directly, not as a difference) is below 3e-15 of the matrix's.
static const int ALG_EIG_MAXN = 2048;
__device__ __forceinline__ double alg_block_sum(double v, double *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < nw; w++) t += red[w];          // same order in every thread: all threads take the same branch later
  return t;
}
__global__ void __launch_bounds__(1024)
alg_jacobi_kernel(const AlgBlk *blks, const double *x, double *a, double *v, double *lab, double *q, int want_v, int *sweeps) {
  extern __shared__ double jac_sm[];
  const AlgBlk B = blks[blockIdx.x];
  const int n = B.n, tid = threadIdx.x, nt = blockDim.x;
  const int m = n + (n & 1), hp = m >> 1;
  double *cs_c = jac_sm, *cs_s = jac_sm + hp;
  int *pp = (int *)(jac_sm + 2 * hp), *pq = pp + hp;
  __shared__ double red[32];
  const double *X = x + B.off;
  double *A = a + B.off, *V = v + B.off, *Q = q ? q + B.off : nullptr;
  double *L = lab + B.voff;
  const long long tot = (long long)n * n;
  for (long long idx = tid; idx < tot; idx += nt) {
    const int i = (int)(idx % n), j = (int)(idx / n);
    A[idx] = 0.5 * (X[idx] + X[j + (long long)i * n]);
    if (want_v) V[idx] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  int sw = 0;
  if (n > 1) {
    double f = 0.0;
    for (long long idx = tid; idx < tot; idx += nt) f += A[idx] * A[idx];
    const double fro2 = alg_block_sum(f, red);
    for (; sw < 30; sw++) {
      double o = 0.0;
      for (long long idx = tid; idx < tot; idx += nt) { const int i = (int)(idx % n), j = (int)(idx / n); if (i != j) o += A[idx] * A[idx]; }
      const double off2 = alg_block_sum(o, red);
      if (!(off2 > 1e-29 * fro2)) break;                        // uniform: every thread holds the same sums
      for (int st = 0; st < m - 1; st++) {
        for (int k = tid; k < hp; k += nt) {
          int i1, i2;
          if (k == 0) { i1 = m - 1; i2 = st % (m - 1); }
          else { i1 = (st + k) % (m - 1); i2 = (st - k + (m - 1)) % (m - 1); }
          const int p = min(i1, i2), qq = max(i1, i2);
          double c = 1.0, sn = 0.0;
          if (qq < n) {
            const double apq = A[p + (long long)qq * n];
            if (apq != 0.0) {
              const double tau = (A[qq + (long long)qq * n] - A[p + (long long)p * n]) / (2.0 * apq);
              const double t = tau == 0.0 ? 1.0 : copysign(1.0, tau) / (fabs(tau) + sqrt(1.0 + tau * tau));
              c = 1.0 / sqrt(1.0 + t * t);
              sn = t * c;
            }
          }
          pp[k] = p; pq[k] = qq; cs_c[k] = c; cs_s[k] = sn;
        }
        __syncthreads();
        for (int idx = tid; idx < hp * hp; idx += nt) {
          const int K = idx % hp, Lq = idx / hp;
          const int p = pp[K], q2 = pq[K], r = pp[Lq], s2 = pq[Lq];
          const double cK = cs_c[K], sK = cs_s[K], cL = cs_c[Lq], sL = cs_s[Lq];
          const bool vq = q2 < n, vs = s2 < n;
          const double b00 = A[p + (long long)r * n];
          const double b10 = vq ? A[q2 + (long long)r * n] : 0.0;
          const double b01 = vs ? A[p + (long long)s2 * n] : 0.0;
          const double b11 = (vq && vs) ? A[q2 + (long long)s2 * n] : 0.0;
          const double t00 = cK * b00 - sK * b10, t10 = sK * b00 + cK * b10;
          const double t01 = cK * b01 - sK * b11, t11 = sK * b01 + cK * b11;
          A[p + (long long)r * n] = t00 * cL - t01 * sL;
          if (vq) A[q2 + (long long)r * n] = t10 * cL - t11 * sL;
          if (vs) A[p + (long long)s2 * n] = t00 * sL + t01 * cL;
          if (vq && vs) A[q2 + (long long)s2 * n] = t10 * sL + t11 * cL;
        }
        if (want_v)
          for (int idx = tid; idx < n * hp; idx += nt) {
            const int i = idx % n, Lq = idx / n;
            const int r = pp[Lq], s2 = pq[Lq];
            if (s2 < n) {
              const double cL = cs_c[Lq], sL = cs_s[Lq];
              const double vr = V[i + (long long)r * n], vs2 = V[i + (long long)s2 * n];
              V[i + (long long)r * n] = cL * vr - sL * vs2;
              V[i + (long long)s2 * n] = sL * vr + cL * vs2;
            }
          }
        __syncthreads();
      }
    }
  }
  if (tid == 0 && sweeps) sweeps[blockIdx.x] = sw;
  // ascending order (ties: original index), eigenvectors follow
  int *rank = pq + hp;                                           // n ints
  for (int i = tid; i < n; i += nt) {
    const double li = A[i + (long long)i * n];
    int rk = 0;
    for (int j = 0; j < n; j++) { const double lj = A[j + (long long)j * n]; rk += (lj < li) || (lj == li && j < i); }
    rank[i] = rk;
    L[rk] = li;
  }
  __syncthreads();
  if (Q)
    for (long long idx = tid; idx < tot; idx += nt) {
      const int r = (int)(idx % n), i = (int)(idx / n);
      Q[r + (long long)rank[i] * n] = V[idx];
    }
}

static int alg_blocks(sb_idx nblk, sb_idx nreal, const sb_idx *n, std::vector<AlgBlk> &blks, long long &lenud, long long &sumn, int &maxn) {
  lenud = 0; sumn = 0; maxn = 0;
  for (sb_idx k = 0; k < nblk; k++) {
    SB_CHECK(n[k] >= 1 && n[k] < 46340, "PSD block order %lld out of range", (long long)n[k]);
    const int c = k >= nreal;
    blks.push_back(AlgBlk{(int)n[k], c, lenud, sumn});
    lenud += (c ? 2 : 1) * n[k] * n[k]; sumn += n[k]; maxn = std::max(maxn, (int)n[k]);
  }
  return 0;
}

}  // namespace sb
using namespace sb;

extern "C" {

// y = vecsym(x, K) on the PSD part (the LP / Lorentz head is copied by the stub); blocks [nreal, nblk) Hermitian.
int sb200_vecsym(sb_idx nblk, sb_idx nreal, const sb_idx *n, const double *x, double *y) {
  SB_TRY(ensure_init());
  std::vector<AlgBlk> blks; long long lenud, sumn; int maxn;
  SB_TRY(alg_blocks(nblk, nreal, n, blks, lenud, sumn, maxn));
  if (lenud == 0) return 0;
  arena_reset();
  cudaStream_t st = ctx().stream;
  AlgBlk *db = arena<AlgBlk>(blks.size()); double *dx = arena<double>((size_t)lenud), *dy = arena<double>((size_t)lenud);
  SB_CHECK(db && dx && dy, "vecsym: out of device memory");
  SB_CUDA(cudaMemcpyAsync(db, blks.data(), sizeof(AlgBlk) * blks.size(), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(dx, x, sizeof(double) * lenud, cudaMemcpyHostToDevice, st));
  vecsym_kernel<<<dim3((unsigned)std::min<long long>(((long long)maxn * maxn + 255) / 256, 1024), (unsigned)nblk), 256, 0, st>>>(db, dx, dy);
  SB_LAUNCH_CHECK_N("vecsym_kernel");
  SB_CUDA(cudaMemcpyAsync(y, dy, sizeof(double) * lenud, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// y = sqrtinv(q, vlab, K): v = PSD part of vlab (sum n doubles)
int sb200_sqrtinv(sb_idx nblk, sb_idx nreal, const sb_idx *n, const double *q, const double *v, double *y) {
  SB_TRY(ensure_init());
  std::vector<AlgBlk> blks; long long lenud, sumn; int maxn;
  SB_TRY(alg_blocks(nblk, nreal, n, blks, lenud, sumn, maxn));
  if (lenud == 0) return 0;
  arena_reset();
  cudaStream_t st = ctx().stream;
  AlgBlk *db = arena<AlgBlk>(blks.size());
  double *dq = arena<double>((size_t)lenud), *dv = arena<double>((size_t)sumn), *dy = arena<double>((size_t)lenud);
  SB_CHECK(db && dq && dv && dy, "sqrtinv: out of device memory");
  SB_CUDA(cudaMemcpyAsync(db, blks.data(), sizeof(AlgBlk) * blks.size(), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(dq, q, sizeof(double) * lenud, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(dv, v, sizeof(double) * sumn, cudaMemcpyHostToDevice, st));
  sqrtinv_kernel<<<dim3((unsigned)std::min<long long>(((long long)maxn * maxn + 255) / 256, 1024), (unsigned)nblk), 256, 0, st>>>(db, dq, dv, dy);
  SB_LAUNCH_CHECK_N("sqrtinv_kernel");
  SB_CUDA(cudaMemcpyAsync(y, dy, sizeof(double) * lenud, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// [q, r] = qrK(x, K), real PSD blocks (qrK.c:263-268); q and r: lenud doubles each.
int sb200_qrK(sb_idx nblk, const sb_idx *n, const double *x, double *q, double *r) {
  SB_TRY(ensure_init());
  std::vector<AlgBlk> blks; long long lenud, sumn; int maxn;
  SB_TRY(alg_blocks(nblk, nblk, n, blks, lenud, sumn, maxn));
  if (lenud == 0) return 0;
  SB_CHECK((size_t)maxn * 8 <= 200 * 1024, "qrK: block order %d too large for the shared-memory reflector", maxn);
  arena_reset();
  cudaStream_t st = ctx().stream;
  AlgBlk *db = arena<AlgBlk>(blks.size()); double *dq = arena<double>((size_t)lenud), *dr = arena<double>((size_t)lenud);
  SB_CHECK(db && dq && dr, "qrK: out of device memory");
  SB_CUDA(cudaMemcpyAsync(db, blks.data(), sizeof(AlgBlk) * blks.size(), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(dr, x, sizeof(double) * lenud, cudaMemcpyHostToDevice, st));
  const size_t shm = sizeof(double) * (size_t)maxn;
  if (shm > 48 * 1024) SB_CUDA(cudaFuncSetAttribute(qrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  qrk_kernel<<<(unsigned)nblk, 256, shm, st>>>(db, dq, dr);
  SB_LAUNCH_CHECK_N("qrk_kernel");
  SB_CUDA(cudaMemcpyAsync(q, dq, sizeof(double) * lenud, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(r, dr, sizeof(double) * lenud, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  return 0;
}


// Batched products Z_k = A_k B_k' on the tile-GEMM engine (one launch): descriptors built per call.
static int alg_gemm(const std::vector<AlgBlk> &blks, int a_tri, int b_tri, const double *dA, const double *dB, double *dC) {
  std::vector<GemmDesc> descs; std::vector<GemmTile> tiles;
  for (size_t k = 0; k < blks.size(); k++) {
    GemmDesc g{}; g.gatherOff = -1; g.alpha = 1.0;
    g.offA = g.offB = g.offC = blks[k].off; g.lda = g.ldb = g.ldc = blks[k].n; g.a_tri = a_tri; g.b_tri = b_tri;
    g.M = g.N = g.K = blks[k].n; g.lower = 0; g.accumulate = 0;
    gemm_add_tiles(tiles, (int)descs.size(), blks[k].n, blks[k].n, false); descs.push_back(g);
  }
  GemmDesc *dd = arena<GemmDesc>(descs.size()); GemmTile *dt = arena<GemmTile>(tiles.size());
  SB_CHECK(dd && dt, "psd algebra: out of device memory");
  cudaStream_t st = ctx().stream;
  SB_CUDA(cudaMemcpyAsync(dd, descs.data(), sizeof(GemmDesc) * descs.size(), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(dt, tiles.data(), sizeof(GemmTile) * tiles.size(), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaStreamSynchronize(st));                 // host vectors go out of scope
  gemm_nt_launch((int)tiles.size(), ctx().sm_count, st, dd, dt, dA, dB, dC, nullptr);
  SB_LAUNCH_CHECK_N("gemm_nt_kernel");
  return 0;
}

// z = psdjmul(x,y,K) (mode 0, psdjmul.m:38-74: Z_k = (X_k Y_k + (X_k Y_k)')/2) and z = triumtriu(x,y,K) (mode 1,
// triumtriu.m:38-73: Z_k = triu(X_k) triu(Y_k), mirrored below the diagonal); real blocks, x/y/z lenud doubles.
int sb200_psdmul(int mode, sb_idx nblk, const sb_idx *n, const double *x, const double *y, double *z) {
  SB_TRY(ensure_init());
  std::vector<AlgBlk> blks; long long lenud, sumn; int maxn;
  SB_TRY(alg_blocks(nblk, nblk, n, blks, lenud, sumn, maxn));
  if (lenud == 0) return 0;
  arena_reset();
  cudaStream_t st = ctx().stream;
  AlgBlk *db = arena<AlgBlk>(blks.size());
  double *dx = arena<double>((size_t)lenud), *dy = arena<double>((size_t)lenud), *dyt = arena<double>((size_t)lenud), *dp = arena<double>((size_t)lenud);
  SB_CHECK(db && dx && dy && dyt && dp, "psdjmul: out of device memory");
  SB_CUDA(cudaMemcpyAsync(db, blks.data(), sizeof(AlgBlk) * blks.size(), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(dx, x, sizeof(double) * lenud, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(dy, y, sizeof(double) * lenud, cudaMemcpyHostToDevice, st));
  const int tp = (maxn + 31) / 32;
  alg_transpose_kernel<<<dim3(tp, tp, (unsigned)nblk), dim3(32, 8), 0, st>>>(db, dy, dyt);
  SB_LAUNCH_CHECK_N("alg_transpose_kernel");
  // P(i,c) = sum_k X(i,k) Yt(c,k); triumtriu: X(i,k) = 0 for k < i, Y(k,c) = 0 for k > c
  SB_TRY(alg_gemm(blks, mode == 1 ? TRI_K_GE_ROW : TRI_NONE, mode == 1 ? TRI_K_LE_ROW : TRI_NONE, dx, dyt, dp));
  alg_sym_kernel<<<dim3((unsigned)std::min<long long>(((long long)maxn * maxn + 255) / 256, 1024), (unsigned)nblk), 256, 0, st>>>(db, dp, dx, mode);
  SB_LAUNCH_CHECK_N("alg_sym_kernel");
  SB_CUDA(cudaMemcpyAsync(z, dx, sizeof(double) * lenud, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// [ux, ispos] = psdfactor(x,K) (psdfactor.m:37-82): lower Cholesky factor of every real block, mirrored; *ispos = 0 and
// the blocks from the first failing one on left zero when a block is not positive definite.
int sb200_psdfactor(sb_idx nblk, const sb_idx *n, const double *x, double *ux, int *ispos) {
  SB_TRY(ensure_init());
  std::vector<AlgBlk> blks; long long lenud, sumn; int maxn;
  SB_TRY(alg_blocks(nblk, nblk, n, blks, lenud, sumn, maxn));
  *ispos = 1;
  if (lenud == 0) return 0;
  arena_reset();
  cudaStream_t st = ctx().stream;
  AlgBlk *db = arena<AlgBlk>(blks.size());
  double *dw = arena<double>((size_t)lenud), *du = arena<double>((size_t)lenud);
  int *dflag = arena<int>(blks.size());
  SB_CHECK(db && dw && du && dflag, "psdfactor: out of device memory");
  SB_CUDA(cudaMemcpyAsync(db, blks.data(), sizeof(AlgBlk) * blks.size(), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(dw, x, sizeof(double) * lenud, cudaMemcpyHostToDevice, st));
  alg_chol_kernel<<<(unsigned)nblk, 256, sizeof(double) * (size_t)maxn, st>>>(db, dw, dflag);
  SB_LAUNCH_CHECK_N("alg_chol_kernel");
  alg_sym_kernel<<<dim3((unsigned)std::min<long long>(((long long)maxn * maxn + 255) / 256, 1024), (unsigned)nblk), 256, 0, st>>>(db, dw, du, 2);
  SB_LAUNCH_CHECK_N("alg_sym_kernel");
  std::vector<int> flag(blks.size());
  SB_CUDA(cudaMemcpyAsync(flag.data(), dflag, sizeof(int) * blks.size(), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(ux, du, sizeof(double) * lenud, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  for (size_t k = 0; k < blks.size(); k++)
    if (flag[k]) { *ispos = 0; memset(ux + blks[k].off, 0, sizeof(double) * (size_t)(lenud - blks[k].off)); break; }
  return 0;
}

// y = psdinvscale(ud,x,K) (psdinvscale.m:37-83): Y_k = T \ (X_k / T'), T = triu(U_k); real blocks.
int sb200_psdinvscale(sb_idx nblk, const sb_idx *n, const double *u, const double *x, double *y) {
  SB_TRY(ensure_init());
  std::vector<AlgBlk> blks; long long lenud, sumn; int maxn;
  SB_TRY(alg_blocks(nblk, nblk, n, blks, lenud, sumn, maxn));
  if (lenud == 0) return 0;
  SB_CHECK(maxn <= 32 * ALG_SOLVE_MAXCH, "psdinvscale: block order %d beyond the register-resident solve (%d)", maxn, 32 * ALG_SOLVE_MAXCH);
  arena_reset();
  cudaStream_t st = ctx().stream;
  AlgBlk *db = arena<AlgBlk>(blks.size());
  const double *du = (const double *)mirror_input(u, sizeof(double) * lenud);
  double *dx = arena<double>((size_t)lenud), *dt = arena<double>((size_t)lenud);
  SB_CHECK(db && du && dx && dt, "psdinvscale: out of device memory");
  SB_CUDA(cudaMemcpyAsync(db, blks.data(), sizeof(AlgBlk) * blks.size(), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(dx, x, sizeof(double) * lenud, cudaMemcpyHostToDevice, st));
  const int tp = (maxn + 31) / 32;
  const dim3 gs((unsigned)((maxn + 7) / 8), (unsigned)nblk);
  // W' = T \ X'  (= (X / T')'), then Y = T \ W
  alg_transpose_kernel<<<dim3(tp, tp, (unsigned)nblk), dim3(32, 8), 0, st>>>(db, dx, dt);
  SB_LAUNCH_CHECK_N("alg_transpose_kernel");
  alg_trsolve_kernel<<<gs, 256, 0, st>>>(db, du, dt);
  SB_LAUNCH_CHECK_N("alg_trsolve_kernel");
  alg_transpose_kernel<<<dim3(tp, tp, (unsigned)nblk), dim3(32, 8), 0, st>>>(db, dt, dx);
  SB_LAUNCH_CHECK_N("alg_transpose_kernel");
  alg_trsolve_kernel<<<gs, 256, 0, st>>>(db, du, dx);
  SB_LAUNCH_CHECK_N("alg_trsolve_kernel");
  SB_CUDA(cudaMemcpyAsync(y, dx, sizeof(double) * lenud, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"

extern "C" {
// [lab, q] = psdeig(x,K) (psdeig.m:40-96, real blocks): eigenvalues of sym(X_k) = (X_k + X_k')/2 in ascending order
// (= 0.5 eig(XX + XX')), q (optional, may be NULL) the orthonormal eigenvectors column by column.  The reference calls the
// host's eig(); here a cyclic Jacobi method (alg_jacobi_kernel), so vectors agree up to sign / rotation inside eigenspaces.
int sb200_psdeig(sb_idx nblk, const sb_idx *n, const double *x, double *lab, double *q) {
  SB_TRY(ensure_init());
  std::vector<AlgBlk> blks; long long lenud, sumn; int maxn;
  SB_TRY(alg_blocks(nblk, nblk, n, blks, lenud, sumn, maxn));
  if (lenud == 0) return 0;
  SB_CHECK(maxn <= ALG_EIG_MAXN, "psdeig: block order %d beyond %d", maxn, ALG_EIG_MAXN);
  arena_reset();
  cudaStream_t st = ctx().stream;
  AlgBlk *db = arena<AlgBlk>(blks.size());
  double *dx = arena<double>((size_t)lenud), *da = arena<double>((size_t)lenud), *dv = arena<double>((size_t)lenud);
  double *dq = q ? arena<double>((size_t)lenud) : nullptr, *dl = arena<double>((size_t)sumn);
  SB_CHECK(db && dx && da && dv && dl && (!q || dq), "psdeig: out of device memory");
  SB_CUDA(cudaMemcpyAsync(db, blks.data(), sizeof(AlgBlk) * blks.size(), cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(dx, x, sizeof(double) * lenud, cudaMemcpyHostToDevice, st));
  const int m = maxn + (maxn & 1), hp = m / 2;
  const size_t shm = sizeof(double) * 2 * hp + sizeof(int) * (2 * (size_t)hp + maxn);
  alg_jacobi_kernel<<<(unsigned)nblk, 1024, shm, st>>>(db, dx, da, dv, dl, dq, q ? 1 : 0, nullptr);
  SB_LAUNCH_CHECK_N("alg_jacobi_kernel");
  SB_CUDA(cudaMemcpyAsync(lab, dl, sizeof(double) * sumn, cudaMemcpyDeviceToHost, st));
  if (q) SB_CUDA(cudaMemcpyAsync(q, dq, sizeof(double) * lenud, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  return 0;
}
}  // extern "C"
