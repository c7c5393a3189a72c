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

}  // extern "C"
