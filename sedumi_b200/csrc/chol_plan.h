// chol_plan.h -- plan structures shared by chol.cu (general supernodal path) and chol_dense.cu
// (single dense supernode fast path).
#pragma once
#include <vector>
#include "sb_internal.h"

namespace sb {

static const int SMALL_N = 128;   // supernodes up to this many columns: one CTA each
static const int NB = 32;         // panel width of the blocked path
static const int UT_R = 64, UT_C = 32;   // update-kernel tile of an ancestor panel

struct Sn {        // one supernode
  int first, n, m; // first column, #columns, #rows of first column (incl. diagonal)
  int lindx;       // offset of its row list in lindx[]
  long long poff;  // offset of its m x n panel in the rect layout (ld = m)
  long long coff;  // offset of its first column in the packed CSC value array (= Ljc[first])
  long long uoff;  // offset of its Schur contribution U = L21 D L21' ((m-n) x (m-n), lower triangle, ld = m-n)
  int cvoff;       // offset of its forward-solve contribution L21 y ((m-n) doubles)
  int pad_;
};
struct Pair {      // update of ancestor J by descendant K
  int K, J;
  int koff;        // first row (index into K's row list) that lies in J's columns
  int mk;          // rows of K from koff to the end
  int ncolup;      // how many of those lie inside J's columns
  int rel;         // offset into rel[]: position of each of those mk rows inside J's row list
};
struct UTile { int J, r0, c0; };

}  // namespace sb

struct sb200_chol_plan {
  int m = 0, nsuper = 0, nlevels = 0;
  long long nnzL = 0, rect = 0;
  sb::Hash128 key;
  uint64_t cache_stamp = 0;      // LRU clock of the host-entry plan cache
  std::vector<sb::Sn> sn;
  std::vector<int> snode, level_of;
  std::vector<std::vector<int>> level_small, level_big;      // supernodes per level
  std::vector<int> level_small_off;                          // offsets into d_level_list
  std::vector<std::vector<sb::UTile>> level_tiles;
  std::vector<int> level_tile_off;
  std::vector<int> pair_beg;                                  // per J: range in pairs[]
  std::vector<std::vector<int>> level_all;                   // all supernodes per level (solves)
  std::vector<int> level_all_off;
  int max_sn_n = 0, max_sn_m = 0;
  // device
  sb::DevBuf<sb::Sn> d_sn;
  sb::DevBuf<sb::Pair> d_pairs;
  sb::DevBuf<int> d_pair_beg, d_rel, d_lindx, d_snode, d_perm, d_Xjc, d_Xir, d_level_list, d_level_all;
  sb::DevBuf<sb::UTile> d_tiles;
  sb::DevBuf<long long> d_Ljc;
  // numeric scratch
  sb::DevBuf<double> d_diagX, d_lb, d_scal, d_vscratch, d_y;
  // contribution blocks of the multi-supernode path: every supernode K forms U_K = L21 D L21' right after it is
  // factored (all supernodes of a level in one launch); its ancestors then only ADD entries of U_K, in list order.
  // The forward solve does the same with the vectors c_K = L21 y_K.
  sb::DevBuf<double> d_U, d_cvec;
  long long utot = 0, ctot = 0;
  // gather lists: for every target (entry of y / entry of an ancestor panel) the contributions that reach it, in
  // descendant order, with the descendant each comes from (sharded runs filter by descendant)
  sb::DevBuf<int> d_pull_ptr, d_pull_src, d_pull_K, d_upd_ptr, d_upd_K;
  sb::DevBuf<long long> d_upd_src;
  bool upd_lists = false;
  int cvec_nrhs = 0, max_mk = 0;
  size_t small_cap = 0;          // doubles of shared memory a factor CTA may use for its panel
  // dense fast path (one supernode spanning the whole matrix): working copy, inverted diagonal
  // blocks (32x32 each, column-major), grid-barrier counter, solve flags
  bool dense_fast = false;
  int npanels = 0;
  sb::DevBuf<double> d_work, d_dinv, d_ys;
  sb::DevBuf<unsigned> d_bar;
  sb::DevBuf<int> d_ready;
  int solve_epoch = 0;
  // ---- subtree sharding: supernodes >= t0 form the replicated "top", the forest below is split over ranks
  struct Shard {
    int world = 1, rank = 0, t0 = 0, top_col0 = 0;
    long long top_rect_off = 0, top_rect_len = 0;
    std::vector<int> owner;                                    // per supernode < t0
    std::vector<std::vector<int>> own_small, own_big, own_all, top_small, top_big, top_all;   // per level
    std::vector<int> own_small_off, own_all_off, top_small_off, top_all_off, own_tile_off, own_tile_cnt, top_tile_off, top_tile_cnt;
    int topA_tile_off = 0, topA_tile_cnt = 0;                  // partial update of the top by owned descendants
    int top_list_off = 0, top_list_cnt = 0;                    // all top supernodes (pull-only forward pass)
    sb::DevBuf<int> d_lists, d_pair_begA, d_pair_begB, d_colmask;
    sb::DevBuf<unsigned char> d_kmaskA, d_kmaskB;              // descendants below the top owned by this rank / inside the top
    sb::DevBuf<sb::UTile> d_tiles;
    sb::DevBuf<sb::Pair> d_pairsA, d_pairsB;
  };
  Shard *shard = nullptr;
  // MEX-level solves: device copy of the last L.L (internal layout) keyed by a hash of its values
  sb::DevBuf<double> d_rect_cache;
  sb::Hash128 Lcache_hash;
  bool Lcache_valid = false;
};

namespace sb {
int dense_factor_prepare(sb200_chol_plan *pl);
int dense_factor(sb200_chol_plan *pl, const double *Xpr, const double *absd, sb200_chol_pars pars, double *rect,
                 double *d, int *flag, double *sval, void (*bounds)(sb200_chol_plan *, const double *, sb200_chol_pars));
int dense_fwsolve(sb200_chol_plan *pl, const double *rect, const double *b, double *y, int nrhs,
                  const double *dscale, const int *flag);
int dense_bwsolve(sb200_chol_plan *pl, const double *rect, const double *b, double *y, int nrhs);
int dense_compute_dinv(sb200_chol_plan *pl, const double *rect);
}
