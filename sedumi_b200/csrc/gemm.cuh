// gemm.cuh -- batched FP64 "NT" GEMM engine used by the PSD-cone kernels.
//
//   C(i,c) (+)= alpha * sum_k A(i,k) * B(c,k)          A: M x K, B: N x K, both column-major
//
// so both operands are read contiguously along the output index (coalesced tile loads).
// A column gather on A (A(i,k) = A[i + gather[k]*lda]) serves getada3's D(:,R) operand;
// triangular hints let tiles skip the k-ranges that are structurally zero (U'U, T'XT).
// One flattened tile list drives a whole batch of independent products in one launch
// (per-PSD-block / per-constraint problems of different sizes), so the grid scales with
// the total work, not with the number of problems.
//
// FP64 on sm_100a has no tcgen05 kind; the tensor path for doubles is the legacy
// mma.sync m8n8k4 DMMA, used by the 64x64 tile kernel below (each warp owns a 32x16
// sub-tile = 4x2 DMMA fragments).  Operands are staged through shared memory.
#pragma once
#include <cuda_runtime.h>

namespace sb {

enum { TRI_NONE = 0, TRI_K_LE_ROW = 1, TRI_K_GE_ROW = 2 };   // nonzero only where k <= row / k >= row

// Operands are addressed as base pointer (kernel argument) + offset (descriptor), so the
// descriptor/tile arrays depend only on the cone structure and live in a plan.
struct GemmDesc {
  long long offA; long long gatherOff; int lda; int a_tri;   // gatherOff < 0: no column gather
  long long offB; int ldb; int b_tri;
  long long offC; int ldc;
  int M, N, K;
  int lower;        // 1: only entries with i >= c are needed (tiles strictly above are skipped)
  int accumulate;   // 0: C = alpha*AB ; 1: C += alpha*AB
  double alpha;
};
struct GemmTile { int prob, ti, tj; };

static const int GT = 64;     // tile edge
static const int GK = 16;     // k-slab

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// 256 threads = 8 warps; warp w owns rows [32*(w&1), +32) x cols [16*(w>>1), +16) of the tile.
static __global__ void __launch_bounds__(256)
gemm_nt_kernel(const GemmDesc *descs, const GemmTile *tiles, const double *baseA, const double *baseB,
               double *baseC, const int *gatherBase) {
  const GemmTile tl = tiles[blockIdx.x];
  const GemmDesc g = descs[tl.prob];
  const double *gA = baseA + g.offA, *gB = baseB + g.offB;
  double *gC = baseC + g.offC;
  const int *gather = (g.gatherOff >= 0) ? gatherBase + g.gatherOff : nullptr;
  const int i0 = tl.ti * GT, c0 = tl.tj * GT;
  __shared__ double As[GK][GT + 4], Bs[GK][GT + 4];
  // k-range that can be nonzero for this tile
  int klo = 0, khi = g.K;
  if (g.a_tri == TRI_K_LE_ROW) khi = min(khi, i0 + GT);
  if (g.a_tri == TRI_K_GE_ROW) klo = max(klo, i0);
  if (g.b_tri == TRI_K_LE_ROW) khi = min(khi, c0 + GT);
  if (g.b_tri == TRI_K_GE_ROW) klo = max(klo, c0);
  klo = (klo / GK) * GK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = (warp & 1) * 32, wc = (warp >> 1) * 16;
  const int qr = lane >> 2, qc = lane & 3;        // fragment coordinates
  double acc[4][2][2];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 2; b++) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
  for (int k0 = klo; k0 < khi; k0 += GK) {
    // stage the slab: element (i, kk) -> As[kk][i]
    for (int idx = threadIdx.x; idx < GT * GK; idx += 256) {
      int i = idx % GT, kk = idx / GT;
      int k = k0 + kk, gi = i0 + i, gc = c0 + i;
      double av = 0.0, bv = 0.0;
      if (k < g.K) {
        if (gi < g.M) {
          bool nz = (g.a_tri == TRI_NONE) || (g.a_tri == TRI_K_LE_ROW ? k <= gi : k >= gi);
          if (nz) av = gA[gi + (long long)(gather ? gather[k] : k) * g.lda];
        }
        if (gc < g.N) {
          bool nz = (g.b_tri == TRI_NONE) || (g.b_tri == TRI_K_LE_ROW ? k <= gc : k >= gc);
          if (nz) bv = gB[gc + (long long)k * g.ldb];
        }
      }
      As[kk][i] = av;
      Bs[kk][i] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int k4 = 0; k4 < GK; k4 += 4) {
      double af[4], bf[2];
#pragma unroll
      for (int a = 0; a < 4; a++) af[a] = As[k4 + qc][wr + 8 * a + qr];     // A frag: row = lane/4, k = lane%4
#pragma unroll
      for (int b = 0; b < 2; b++) bf[b] = Bs[k4 + qc][wc + 8 * b + qr];     // B frag: k = lane%4, col = lane/4
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
    __syncthreads();
  }
  // C fragment: row = lane/4, cols = 2*(lane%4) + {0,1}
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 2; b++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        int gi = i0 + wr + 8 * a + qr, gc = c0 + wc + 8 * b + 2 * qc + e;
        if (gi < g.M && gc < g.N && (!g.lower || gi >= gc)) {
          double *p = gC + gi + (long long)gc * g.ldc;
          double v = g.alpha * acc[a][b][e];
          *p = g.accumulate ? (*p + v) : v;
        }
      }
}

// Host helper: append the tiles of one problem to a tile list.
template <typename Vec>
inline void gemm_add_tiles(Vec &tiles, int prob, int M, int N, bool lower) {
  for (int tj = 0; tj * GT < N; tj++)
    for (int ti = 0; ti * GT < M; ti++) {
      if (lower && (ti + 1) * GT - 1 < tj * GT) continue;
      tiles.push_back(GemmTile{prob, ti, tj});
    }
}

}  // namespace sb
