// psd.cu -- per-PSD-block dense algebra: invcholfac (D = U'U) and psdscale (Y = T'XT).
//
// Reference semantics:
//   invcholfac.c:122-141  per block: Z = triu(U)'*triu(U) (utmulx, triuaux.c:175-186),
//                         symmetrised, then Y(perm,perm) = Z (invmatperm, triuaux.c:61-72)
//   psdscale.m:76-110     per block: T = tril(U) (transp=0) or triu(U) (transp=1) of the
//                         stored n x n array, X optionally X(perm,perm) before (transp=0)
//                         or the result scattered to (perm,perm) after (transp=1); Y = T'*X*T
//
// GPU design: every block is a set of 64x64 tiles in ONE batched launch of the DMMA tile
// GEMM (gemm.cuh).  Both operations are written as "NT" products over Tt = T' so that all
// tile loads are contiguous:
//   invcholfac:  Z(i,j)  = sum_k Tt(i,k) Tt(j,k)            (Tt = triu(U)', lower triangular)
//   psdscale:    Wt(c,i) = sum_k Tt(c,k) X(i,k);  Y(i,c) = sum_k Tt(i,k) Wt(c,k)
// and the triangular k-ranges are skipped per tile.  Descriptors and tile lists depend only
// on K.s, so they live in a cached plan; per call only the block data moves.
#include <map>
#include "gemm.cuh"
#include "sb_internal.h"

struct sb200_psd_plan {
  int nblk = 0;
  long long lenud = 0, sumn = 0;
  int maxn = 0;
  std::vector<int> n;
  std::vector<long long> off;      // offset of block k in a lenud vector
  std::vector<int> poff;           // offset of block k in a perm vector
  // device
  sb::DevBuf<int> d_n, d_poff;
  sb::DevBuf<long long> d_off;
  sb::DevBuf<sb::GemmDesc> d_desc_ichol, d_desc_s1_lo, d_desc_s1_up, d_desc_s2_lo, d_desc_s2_up, d_desc_full, d_desc_lower;
  sb::DevBuf<int> d_qcol_blk, d_qcol_j0;     // column groups for the Householder accumulation
  int nqgroups = 0;
  sb::DevBuf<sb::GemmTile> d_tiles_full, d_tiles_lower;
  int ntiles_full = 0, ntiles_lower = 0;
  // workspaces (lenud doubles each)
  sb::DevBuf<double> d_Tt, d_Wt, d_Xp, d_Y;
  sb::DevBuf<int> d_perm;
};

namespace sb {

// blockIdx.y = PSD block; Tt(c,k) = mask ? U[k + c*n] : 0 with mask (k<=c) for upper=1, (k>=c) for upper=0
__global__ void tri_transpose_kernel(const int *ns, const long long *offs, const double *u, double *Tt, int upper) {
  const int n = ns[blockIdx.y];
  const double *U = u + offs[blockIdx.y];
  double *T = Tt + offs[blockIdx.y];
  __shared__ double tile[32][33];
  const int tilesPerDim = (n + 31) / 32;
  for (int t = blockIdx.x; t < tilesPerDim * tilesPerDim; t += gridDim.x) {
    int k0 = (t % tilesPerDim) * 32, c0 = (t / tilesPerDim) * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      int k = k0 + threadIdx.x, c = c0 + j;
      tile[j][threadIdx.x] = (k < n && c < n) ? U[k + (long long)c * n] : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      int c = c0 + threadIdx.x, k = k0 + j;
      if (c < n && k < n) {
        bool keep = upper ? (k <= c) : (k >= c);
        T[c + (long long)k * n] = keep ? tile[threadIdx.x][j] : 0.0;
      }
    }
    __syncthreads();
  }
}

// out(i,j) = in(p[i], p[j])  (gather=1)   or   out(p[i], p[j]) = in(i,j)  (gather=0); sym=1 reads in(max,min)
__global__ void perm_block_kernel(const int *ns, const long long *offs, const int *poffs, const int *perm,
                                  const double *in, double *out, int gather, int sym) {
  const int n = ns[blockIdx.y];
  const long long off = offs[blockIdx.y];
  const int *p = perm ? perm + poffs[blockIdx.y] : nullptr;
  const long long tot = (long long)n * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < tot; idx += (long long)gridDim.x * blockDim.x) {
    int i = (int)(idx % n), j = (int)(idx / n);
    int pi = p ? p[i] : i, pj = p ? p[j] : j;
    if (gather) out[off + idx] = in[off + pi + (long long)pj * n];
    else {
      double v = sym ? in[off + max(i, j) + (long long)min(i, j) * n] : in[off + idx];
      out[off + pi + (long long)pj * n] = v;
    }
  }
}


// ---------------------------------------------------------------- Householder frames (psdframeit / psdinvjmul)
// frms block k: n x n, column c (c < n-1) holds reflector vector c_c in rows c..n-1, last column holds
// beta_0..beta_{n-2};  H_c = I - c_c c_c'/beta_c ;  Qb = H_0 H_1 ... H_{n-2}   (reflect.c:203-215, qrK.c:86-122).
// Column j of Qb is H_0...H_{min(j,n-2)} e_j and columns are independent: one warp per column,
// 8 columns per CTA, the reflectors streamed from L2.  Q is written column-major.
__global__ void __launch_bounds__(256)
householder_q_kernel(const int *grp_blk, const int *grp_j0, const int *ns, const long long *offs,
                     const double *frms, double *Q) {
  const int k = grp_blk[blockIdx.x];
  const int n = ns[k];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = grp_j0[blockIdx.x] + warp;
  if (j >= n) return;
  const double *F = frms + offs[k];
  const double *beta = F + (long long)n * n - n;
  double *q = Q + offs[k] + (long long)j * n;
  for (int i = lane; i < n; i += 32) q[i] = (i == j) ? 1.0 : 0.0;
  __syncwarp();
  for (int c = min(j, n - 2); c >= 0; c--) {
    const double *v = F + (long long)c * n;           // v[i], i = c..n-1
    double t = 0.0;
    for (int i = c + lane; i < n; i += 32) t += v[i] * q[i];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    const double a = t / (-beta[c]);                  // elqxq is called with -beta (reflect.c:214)
    for (int i = c + lane; i < n; i += 32) q[i] += a * v[i];
    __syncwarp();
  }
}

// out = transpose(in) per block, optionally scaling row k of the result's k-index:  out(c,k) = in(k,c) * (lab ? lab[k] : 1)
__global__ void transpose_scale_kernel(const int *ns, const long long *offs, const int *poffs, const double *in,
                                       const double *lab, double *out) {
  const int n = ns[blockIdx.y];
  const double *A = in + offs[blockIdx.y];
  double *B = out + offs[blockIdx.y];
  const double *lb = lab ? lab + poffs[blockIdx.y] : nullptr;
  __shared__ double tile[32][33];
  const int tpd = (n + 31) / 32;
  for (int t = blockIdx.x; t < tpd * tpd; t += gridDim.x) {
    int k0 = (t % tpd) * 32, c0 = (t / tpd) * 32;
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
      int k = k0 + threadIdx.x, c = c0 + jj;
      tile[jj][threadIdx.x] = (k < n && c < n) ? A[k + (long long)c * n] : 0.0;
    }
    __syncthreads();
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
      int c = c0 + threadIdx.x, k = k0 + jj;
      if (c < n && k < n) B[c + (long long)k * n] = tile[threadIdx.x][jj] * (lb ? lb[k] : 1.0);
    }
    __syncthreads();
  }
}

// mode 0: mirror the lower triangle up (tril2sym); mode 1: out = symmetric matrix built from tril(in);
// mode 2: in-place jdiv on the lower triangle then mirror: m(i,j) *= 2/(x_i+x_j), m(j,j) /= x_j (psdinvjmul.c:69-84)
__global__ void sym_ops_kernel(const int *ns, const long long *offs, const int *poffs, const double *in, double *out,
                               const double *x, int mode) {
  const int n = ns[blockIdx.y];
  const long long off = offs[blockIdx.y];
  const double *xx = x ? x + poffs[blockIdx.y] : nullptr;
  const long long tot = (long long)n * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < tot; idx += (long long)gridDim.x * blockDim.x) {
    int i = (int)(idx % n), j = (int)(idx / n);
    if (mode == 1) { out[off + idx] = in[off + max(i, j) + (long long)min(i, j) * n]; continue; }
    if (i < j) continue;                              // the lower-triangle thread owns the pair (i,j),(j,i)
    double v = in[off + i + (long long)j * n];
    if (mode == 2) v = (i == j) ? v / xx[j] : v * (2.0 / (xx[i] + xx[j]));
    out[off + i + (long long)j * n] = v;
    if (i != j) out[off + j + (long long)i * n] = v;
  }
}

// ---------------------------------------------------------------- fused psdscale for small blocks
// One CTA per PSD block, everything in shared memory: T (masked triangle of U), X (optionally
// X(perm,perm)), W = X*T, then Y = T'*W straight to global (optionally scattered to (perm,perm)).
// Replaces 4-5 launches per call by one when max n_k <= PSD_SMALL_MAX (psdscale.m:76-110).
static const int PSD_SMALL_MAX = 96;
static const int PSD_CS = 16;        // output columns per CTA
// Y(:,c) = T' * (X * T(:,c)) is independent per output column c, so a block is split over
// ceil(n/16) CTAs (blockIdx.y) that each keep T and X in shared memory and own 16 columns of W and Y:
// no inter-CTA synchronisation, 4-5x more SMs busy than one CTA per block.
// 512 threads = 32 (rows tx) x 16 (columns ty); thread owns rows {tx+32a, a<NR} of column ty.
template <int NR>
__global__ void __launch_bounds__(512)
psdscale_small_kernel(const int *ns, const long long *offs, const int *poffs, const double *u, const int *perm,
                      const double *x, int transp, double *y) {
  extern __shared__ double sm[];
  constexpr int NP = 32 * NR, LD = NP + 1;
  const int n = ns[blockIdx.x];
  const int c0 = blockIdx.y * PSD_CS;
  if (c0 >= n) return;
  double *T = sm, *X = sm + NP * LD, *W = X + NP * LD;      // W: NP x PSD_CS (ld = LD)
  const double *U = u + offs[blockIdx.x], *Xg = x + offs[blockIdx.x];
  double *Yg = y + offs[blockIdx.x];
  const int *p = perm ? perm + poffs[blockIdx.x] : nullptr;
  const bool prep = p && !transp, postp = p && transp;
  for (int idx = threadIdx.x; idx < NP * NP; idx += blockDim.x) {
    const int i = idx % NP, k = idx / NP;
    double tv = 0.0, xv = 0.0;
    if (i < n && k < n) {
      const bool keep = transp ? (i <= k) : (i >= k);          // triu : tril  of the stored array
      tv = keep ? U[i + (long long)k * n] : 0.0;
      xv = prep ? Xg[p[i] + (long long)p[k] * n] : Xg[i + (long long)k * n];
    }
    T[i + k * LD] = tv;
    X[i + k * LD] = xv;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // ty < 16
  const int c = c0 + ty;
  {  // W(i,c) = sum_k X(i,k) T(k,c)
    double acc[NR] = {};
    if (c < n) {
#pragma unroll 4
      for (int k = 0; k < n; k++) {
        const double bv = T[k + c * LD];
#pragma unroll
        for (int a = 0; a < NR; a++) acc[a] += X[tx + 32 * a + k * LD] * bv;
      }
    }
#pragma unroll
    for (int a = 0; a < NR; a++) W[tx + 32 * a + ty * LD] = acc[a];
  }
  __syncthreads();
  {  // Y(i,c) = sum_k T(k,i) W(k,c)
    double acc[NR] = {};
    if (c < n) {
#pragma unroll 4
      for (int k = 0; k < n; k++) {
        const double bv = W[k + ty * LD];
#pragma unroll
        for (int a = 0; a < NR; a++) acc[a] += T[k + (tx + 32 * a) * LD] * bv;
      }
#pragma unroll
      for (int a = 0; a < NR; a++) {
        const int i = tx + 32 * a;
        if (i < n) {
          if (postp) Yg[p[i] + (long long)p[c] * n] = acc[a];
          else Yg[i + (long long)c * n] = acc[a];
        }
      }
    }
  }
}

static std::map<uint64_t, sb200_psd_plan *> g_psd_plans;

}  // namespace sb

using namespace sb;

static int psd_build(sb200_psd_plan *pl, sb_idx nblk, const sb_idx *n) {
  pl->nblk = (int)nblk;
  long long off = 0; int po = 0;
  for (sb_idx k = 0; k < nblk; k++) {
    SB_CHECK(n[k] >= 1 && n[k] < 46340, "PSD block order %lld out of range", (long long)n[k]);
    pl->n.push_back((int)n[k]); pl->off.push_back(off); pl->poff.push_back(po);
    off += n[k] * n[k]; po += (int)n[k];
    pl->maxn = std::max(pl->maxn, (int)n[k]);
  }
  pl->lenud = off; pl->sumn = po;
  std::vector<GemmDesc> ichol, s1lo, s1up, s2lo, s2up, dfull, dlower;
  std::vector<int> qblk, qj0;
  std::vector<GemmTile> tfull, tlower;
  for (int k = 0; k < pl->nblk; k++) {
    int nk = pl->n[k]; long long o = pl->off[k];
    GemmDesc g{};
    g.gatherOff = -1; g.lda = g.ldb = g.ldc = nk; g.M = g.N = g.K = nk; g.alpha = 1.0; g.accumulate = 0;
    g.offA = g.offB = g.offC = o;
    // invcholfac: A = B = Tt (k <= row), lower only
    GemmDesc a = g; a.a_tri = TRI_K_LE_ROW; a.b_tri = TRI_K_LE_ROW; a.lower = 1; ichol.push_back(a);
    // psdscale stage 1: Wt = Tt * X'   (A = Tt, B = X)
    GemmDesc b = g; b.lower = 0; b.b_tri = TRI_NONE;
    b.a_tri = TRI_K_GE_ROW; s1lo.push_back(b);          // T = tril(U): Tt(c,k) nonzero for k >= c
    b.a_tri = TRI_K_LE_ROW; s1up.push_back(b);          // T = triu(U)
    // stage 2: Y = Tt * Wt'
    b.a_tri = TRI_K_GE_ROW; s2lo.push_back(b);
    b.a_tri = TRI_K_LE_ROW; s2up.push_back(b);
    GemmDesc f = g; f.a_tri = f.b_tri = TRI_NONE; f.lower = 0; dfull.push_back(f);
    f.lower = 1; dlower.push_back(f);
    for (int j0 = 0; j0 < nk; j0 += 8) { qblk.push_back(k); qj0.push_back(j0); }
    gemm_add_tiles(tfull, k, nk, nk, false);
    gemm_add_tiles(tlower, k, nk, nk, true);
  }
  pl->ntiles_full = (int)tfull.size(); pl->ntiles_lower = (int)tlower.size();
  SB_TRY(pl->d_n.upload(pl->n)); SB_TRY(pl->d_off.upload(pl->off)); SB_TRY(pl->d_poff.upload(pl->poff));
  SB_TRY(pl->d_desc_ichol.upload(ichol));
  SB_TRY(pl->d_desc_s1_lo.upload(s1lo)); SB_TRY(pl->d_desc_s1_up.upload(s1up));
  SB_TRY(pl->d_desc_s2_lo.upload(s2lo)); SB_TRY(pl->d_desc_s2_up.upload(s2up));
  SB_TRY(pl->d_desc_full.upload(dfull)); SB_TRY(pl->d_desc_lower.upload(dlower));
  pl->nqgroups = (int)qblk.size();
  SB_TRY(pl->d_qcol_blk.upload(qblk)); SB_TRY(pl->d_qcol_j0.upload(qj0));
  SB_TRY(pl->d_tiles_full.upload(tfull)); SB_TRY(pl->d_tiles_lower.upload(tlower));
  SB_TRY(pl->d_Tt.alloc((size_t)pl->lenud)); SB_TRY(pl->d_Wt.alloc((size_t)pl->lenud));
  SB_TRY(pl->d_Xp.alloc((size_t)pl->lenud)); SB_TRY(pl->d_Y.alloc((size_t)pl->lenud));
  SB_TRY(pl->d_perm.alloc((size_t)std::max<long long>(pl->sumn, 1)));
  SB_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}

extern "C" {

int sb200_psd_plan_get(sb200_psd_plan **plan, sb_idx nblk, const sb_idx *n) {
  SB_TRY(ensure_init());
  uint64_t key = fnv1a(n, sizeof(sb_idx) * nblk, fnv1a(&nblk, sizeof nblk));
  auto it = g_psd_plans.find(key);
  if (it != g_psd_plans.end()) { *plan = it->second; return 0; }
  sb200_psd_plan *pl = new sb200_psd_plan();
  int rc = psd_build(pl, nblk, n);
  if (rc) { delete pl; return rc; }
  g_psd_plans[key] = pl;
  *plan = pl;
  return 0;
}
sb_idx sb200_psd_plan_lenud(const sb200_psd_plan *pl) { return pl->lenud; }
sb_idx sb200_psd_plan_sumn(const sb200_psd_plan *pl) { return pl->sumn; }

static dim3 blk_grid(const sb200_psd_plan *pl, int per) {
  long long t = (long long)pl->maxn * pl->maxn;
  int gx = (int)std::min<long long>((t + per - 1) / per, 1024);
  return dim3(std::max(gx, 1), pl->nblk);
}

// y = invcholfac(u, K, perm): perm_dev is int32 0-based (length sum n_k) or NULL.
int sb200_invcholfac_dev(sb200_psd_plan *pl, const double *u_dev, const int *perm_dev, double *y_dev) {
  SB_TRY(ensure_init());
  if (pl->nblk == 0) return 0;
  cudaStream_t st = ctx().stream;
  tri_transpose_kernel<<<dim3(std::min(1024, ((pl->maxn + 31) / 32) * ((pl->maxn + 31) / 32)), pl->nblk), dim3(32, 8), 0, st>>>(
      pl->d_n.p, pl->d_off.p, u_dev, pl->d_Tt.p, 1);
  SB_LAUNCH_CHECK_N("tri_transpose_kernel");
  gemm_nt_launch(pl->ntiles_lower, ctx().sm_count, st, pl->d_desc_ichol.p, pl->d_tiles_lower.p, pl->d_Tt.p, pl->d_Tt.p,
                                                   pl->d_Wt.p, nullptr);
  SB_LAUNCH_CHECK_N("gemm_nt_kernel");
  perm_block_kernel<<<blk_grid(pl, 256), 256, 0, st>>>(pl->d_n.p, pl->d_off.p, pl->d_poff.p, perm_dev, pl->d_Wt.p, y_dev, 0, 1);
  SB_LAUNCH_CHECK_N("perm_block_kernel");
  return 0;
}

// y = psdscale(ud, x, K, transp): x_dev/y_dev are the PSD parts (lenud).  perm_dev as above
// (NULL when ud.perm is empty or ud is a plain vector).
int sb200_psdscale_dev(sb200_psd_plan *pl, const double *u_dev, const int *perm_dev, const double *x_dev,
                       int transp, double *y_dev) {
  SB_TRY(ensure_init());
  if (pl->nblk == 0) return 0;
  cudaStream_t st = ctx().stream;
  if (pl->maxn <= PSD_SMALL_MAX) {
    const int NR = (pl->maxn + 31) / 32, NP = 32 * NR;
    const size_t shm = sizeof(double) * (2 * (size_t)NP + PSD_CS) * (NP + 1);
    auto launch = [&](auto kern) -> int {
      if (shm > 48 * 1024) SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
      kern<<<dim3(pl->nblk, (pl->maxn + PSD_CS - 1) / PSD_CS), 512, shm, st>>>(pl->d_n.p, pl->d_off.p, pl->d_poff.p, u_dev, perm_dev, x_dev, transp, y_dev);
      return 0;
    };
    if (NR == 1) SB_TRY(launch(psdscale_small_kernel<1>));
    else if (NR == 2) SB_TRY(launch(psdscale_small_kernel<2>));
    else SB_TRY(launch(psdscale_small_kernel<3>));
    SB_LAUNCH_CHECK_N("psdscale_small_kernel");
    return 0;
  }
  tri_transpose_kernel<<<dim3(std::min(1024, ((pl->maxn + 31) / 32) * ((pl->maxn + 31) / 32)), pl->nblk), dim3(32, 8), 0, st>>>(
      pl->d_n.p, pl->d_off.p, u_dev, pl->d_Tt.p, transp ? 1 : 0);
  SB_LAUNCH_CHECK_N("tri_transpose_kernel");
  const double *xs = x_dev;
  if (perm_dev && !transp) {          // prep: X(perm,perm)   (psdscale.m:94-99)
    perm_block_kernel<<<blk_grid(pl, 256), 256, 0, st>>>(pl->d_n.p, pl->d_off.p, pl->d_poff.p, perm_dev, x_dev, pl->d_Xp.p, 1, 0);
    SB_LAUNCH_CHECK_N("perm_block_kernel");
    xs = pl->d_Xp.p;
  }
  gemm_nt_launch(pl->ntiles_full, ctx().sm_count, st, (transp ? pl->d_desc_s1_up : pl->d_desc_s1_lo).p, pl->d_tiles_full.p,
                                                  pl->d_Tt.p, xs, pl->d_Wt.p, nullptr);
  SB_LAUNCH_CHECK_N("gemm_nt_kernel");
  bool postp = perm_dev && transp;
  gemm_nt_launch(pl->ntiles_full, ctx().sm_count, st, (transp ? pl->d_desc_s2_up : pl->d_desc_s2_lo).p, pl->d_tiles_full.p,
                                                  pl->d_Tt.p, pl->d_Wt.p, postp ? pl->d_Y.p : y_dev, nullptr);
  SB_LAUNCH_CHECK_N("gemm_nt_kernel");
  if (postp) {                         // XX(PP,PP) = XX   (psdscale.m:104-109)
    perm_block_kernel<<<blk_grid(pl, 256), 256, 0, st>>>(pl->d_n.p, pl->d_off.p, pl->d_poff.p, perm_dev, pl->d_Y.p, y_dev, 0, 0);
    SB_LAUNCH_CHECK_N("perm_block_kernel");
  }
  return 0;
}

static int upload_perm(sb200_psd_plan *pl, const sb_idx *perm, const int **out) {
  *out = nullptr;
  if (!perm) return 0;
  std::vector<int> p32((size_t)pl->sumn);
  for (int k = 0; k < pl->nblk; k++)
    for (int i = 0; i < pl->n[k]; i++) {
      sb_idx v = perm[pl->poff[k] + i];
      SB_CHECK(v >= 0 && v < pl->n[k], "perm entry out of range in PSD block %d", k);
      p32[pl->poff[k] + i] = (int)v;
    }
  SB_CUDA(cudaMemcpyAsync(pl->d_perm.p, p32.data(), sizeof(int) * p32.size(), cudaMemcpyHostToDevice, ctx().stream));
  SB_CUDA(cudaStreamSynchronize(ctx().stream));
  *out = pl->d_perm.p;
  return 0;
}

// Host entries.  perm: 0-based within each block (NULL = none).
int sb200_invcholfac(sb_idx nblk, const sb_idx *n, const double *u, const sb_idx *perm, double *y) {
  sb200_psd_plan *pl = nullptr;
  SB_TRY(sb200_psd_plan_get(&pl, nblk, n));
  if (pl->lenud == 0) return 0;
  arena_reset();
  double *du = arena<double>((size_t)pl->lenud), *dy = arena<double>((size_t)pl->lenud);
  SB_CHECK(du && dy, "invcholfac: out of device memory");
  const int *dperm;
  SB_TRY(upload_perm(pl, perm, &dperm));
  SB_CUDA(cudaMemcpyAsync(du, u, sizeof(double) * pl->lenud, cudaMemcpyHostToDevice, ctx().stream));
  SB_TRY(sb200_invcholfac_dev(pl, du, dperm, dy));
  SB_CUDA(cudaMemcpyAsync(y, dy, sizeof(double) * pl->lenud, cudaMemcpyDeviceToHost, ctx().stream));
  SB_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}

int sb200_psdscale(sb_idx nblk, const sb_idx *n, const double *u, const sb_idx *perm, const double *x, int transp, double *y) {
  sb200_psd_plan *pl = nullptr;
  SB_TRY(sb200_psd_plan_get(&pl, nblk, n));
  if (pl->lenud == 0) return 0;
  arena_reset();
  double *du = arena<double>((size_t)pl->lenud), *dx = arena<double>((size_t)pl->lenud), *dy = arena<double>((size_t)pl->lenud);
  SB_CHECK(du && dx && dy, "psdscale: out of device memory");
  const int *dperm;
  SB_TRY(upload_perm(pl, perm, &dperm));
  SB_CUDA(cudaMemcpyAsync(du, u, sizeof(double) * pl->lenud, cudaMemcpyHostToDevice, ctx().stream));
  SB_CUDA(cudaMemcpyAsync(dx, x, sizeof(double) * pl->lenud, cudaMemcpyHostToDevice, ctx().stream));
  SB_TRY(sb200_psdscale_dev(pl, du, dperm, dx, transp, dy));
  SB_CUDA(cudaMemcpyAsync(y, dy, sizeof(double) * pl->lenud, cudaMemcpyDeviceToHost, ctx().stream));
  SB_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}

// ---- Householder-frame operations.  frms_dev: lenud doubles (real blocks); lab/xlab: sum(n_k) doubles.
static int build_q(sb200_psd_plan *pl, const double *frms_dev) {   // Q -> d_Tt, Q' -> d_Wt
  cudaStream_t st = ctx().stream;
  householder_q_kernel<<<pl->nqgroups, 256, 0, st>>>(pl->d_qcol_blk.p, pl->d_qcol_j0.p, pl->d_n.p, pl->d_off.p, frms_dev, pl->d_Tt.p);
  SB_LAUNCH_CHECK_N("householder_q_kernel");
  int tp = (pl->maxn + 31) / 32;
  transpose_scale_kernel<<<dim3(std::min(1024, tp * tp), pl->nblk), dim3(32, 8), 0, st>>>(pl->d_n.p, pl->d_off.p, pl->d_poff.p, pl->d_Tt.p, nullptr, pl->d_Wt.p);
  SB_LAUNCH_CHECK_N("transpose_scale_kernel");
  return 0;
}

// x = psdframeit(lab, frms, K):  X_k = Qb' diag(lab_k) Qb   (psdframeit.c:65-99)
int sb200_psdframeit_dev(sb200_psd_plan *pl, const double *lab_dev, const double *frms_dev, double *x_dev) {
  SB_TRY(ensure_init());
  if (pl->nblk == 0) return 0;
  cudaStream_t st = ctx().stream;
  SB_TRY(build_q(pl, frms_dev));
  int tp = (pl->maxn + 31) / 32;
  // B(c,k) = lab_k * Qb(k,c): transpose of Q with the k-index scaled -> d_Xp
  transpose_scale_kernel<<<dim3(std::min(1024, tp * tp), pl->nblk), dim3(32, 8), 0, st>>>(pl->d_n.p, pl->d_off.p, pl->d_poff.p, pl->d_Tt.p, lab_dev, pl->d_Xp.p);
  SB_LAUNCH_CHECK_N("transpose_scale_kernel");
  gemm_nt_launch(pl->ntiles_lower, ctx().sm_count, st, pl->d_desc_lower.p, pl->d_tiles_lower.p, pl->d_Wt.p, pl->d_Xp.p, x_dev, nullptr);
  SB_LAUNCH_CHECK_N("gemm_nt_kernel");
  sym_ops_kernel<<<blk_grid(pl, 256), 256, 0, st>>>(pl->d_n.p, pl->d_off.p, pl->d_poff.p, x_dev, x_dev, nullptr, 0);
  SB_LAUNCH_CHECK_N("sym_ops_kernel");
  return 0;
}

// z = psdinvjmul(xlab, xfrm, y, K):  solve X Z + Z X = 2 Y in the eigenbasis of X   (psdinvjmul.c:101-157)
int sb200_psdinvjmul_dev(sb200_psd_plan *pl, const double *xlab_dev, const double *frms_dev, const double *y_dev, double *z_dev) {
  SB_TRY(ensure_init());
  if (pl->nblk == 0) return 0;
  cudaStream_t st = ctx().stream;
  SB_TRY(build_q(pl, frms_dev));                                      // Q = d_Tt, Q' = d_Wt
  sym_ops_kernel<<<blk_grid(pl, 256), 256, 0, st>>>(pl->d_n.p, pl->d_off.p, pl->d_poff.p, y_dev, pl->d_Y.p, nullptr, 1);   // Ys from tril(Y)
  SB_LAUNCH_CHECK_N("sym_ops_kernel");
  // P = Q Ys           (A = Q, B = Ys symmetric)            -> d_Xp
  gemm_nt_launch(pl->ntiles_full, ctx().sm_count, st, pl->d_desc_full.p, pl->d_tiles_full.p, pl->d_Tt.p, pl->d_Y.p, pl->d_Xp.p, nullptr);
  SB_LAUNCH_CHECK_N("gemm_nt_kernel");
  // M = P Q' (lower)   (A = P, B = Q)                       -> d_Y
  gemm_nt_launch(pl->ntiles_lower, ctx().sm_count, st, pl->d_desc_lower.p, pl->d_tiles_lower.p, pl->d_Xp.p, pl->d_Tt.p, pl->d_Y.p, nullptr);
  SB_LAUNCH_CHECK_N("gemm_nt_kernel");
  sym_ops_kernel<<<blk_grid(pl, 256), 256, 0, st>>>(pl->d_n.p, pl->d_off.p, pl->d_poff.p, pl->d_Y.p, pl->d_Y.p, xlab_dev, 2);
  SB_LAUNCH_CHECK_N("sym_ops_kernel");
  // R = Q' M           (A = Q', B = M symmetric)            -> d_Xp
  gemm_nt_launch(pl->ntiles_full, ctx().sm_count, st, pl->d_desc_full.p, pl->d_tiles_full.p, pl->d_Wt.p, pl->d_Y.p, pl->d_Xp.p, nullptr);
  SB_LAUNCH_CHECK_N("gemm_nt_kernel");
  // Z = R Q (lower)    (A = R, B = Q')                      -> z
  gemm_nt_launch(pl->ntiles_lower, ctx().sm_count, st, pl->d_desc_lower.p, pl->d_tiles_lower.p, pl->d_Xp.p, pl->d_Wt.p, z_dev, nullptr);
  SB_LAUNCH_CHECK_N("gemm_nt_kernel");
  sym_ops_kernel<<<blk_grid(pl, 256), 256, 0, st>>>(pl->d_n.p, pl->d_off.p, pl->d_poff.p, z_dev, z_dev, nullptr, 0);
  SB_LAUNCH_CHECK_N("sym_ops_kernel");
  return 0;
}

int sb200_psdframeit(sb_idx nblk, const sb_idx *n, const double *lab, const double *frms, double *x) {
  sb200_psd_plan *pl = nullptr;
  SB_TRY(sb200_psd_plan_get(&pl, nblk, n));
  if (pl->lenud == 0) return 0;
  arena_reset();
  double *dl = arena<double>((size_t)pl->sumn), *df = arena<double>((size_t)pl->lenud), *dx = arena<double>((size_t)pl->lenud);
  SB_CHECK(dl && df && dx, "psdframeit: out of device memory");
  cudaStream_t st = ctx().stream;
  SB_CUDA(cudaMemcpyAsync(dl, lab, sizeof(double) * pl->sumn, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(df, frms, sizeof(double) * pl->lenud, cudaMemcpyHostToDevice, st));
  SB_TRY(sb200_psdframeit_dev(pl, dl, df, dx));
  SB_CUDA(cudaMemcpyAsync(x, dx, sizeof(double) * pl->lenud, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int sb200_psdinvjmul(sb_idx nblk, const sb_idx *n, const double *xlab, const double *frms, const double *y, double *z) {
  sb200_psd_plan *pl = nullptr;
  SB_TRY(sb200_psd_plan_get(&pl, nblk, n));
  if (pl->lenud == 0) return 0;
  arena_reset();
  double *dl = arena<double>((size_t)pl->sumn), *df = arena<double>((size_t)pl->lenud), *dy = arena<double>((size_t)pl->lenud),
         *dz = arena<double>((size_t)pl->lenud);
  SB_CHECK(dl && df && dy && dz, "psdinvjmul: out of device memory");
  cudaStream_t st = ctx().stream;
  SB_CUDA(cudaMemcpyAsync(dl, xlab, sizeof(double) * pl->sumn, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(df, frms, sizeof(double) * pl->lenud, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaMemcpyAsync(dy, y, sizeof(double) * pl->lenud, cudaMemcpyHostToDevice, st));
  SB_TRY(sb200_psdinvjmul_dev(pl, dl, df, dy, dz));
  SB_CUDA(cudaMemcpyAsync(z, dz, sizeof(double) * pl->lenud, cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
