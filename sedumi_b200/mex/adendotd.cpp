// Ad = adendotd(dense,d,sparAd,Ablk,blkstart)   (ai[k]+Adeni[k])'*d[k] for the dense Lorentz blocks
// (adendotd.c:44-52 signature, :134-245 mexFunction; called from getDAtm.m:45)
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 5, "adendotd requires more input arguments");
  MEX_REQUIRE(nlhs <= 1, "adendotd produces less output arguments");
  const mxArray *DENSE = prhs[0], *D = prhs[1], *ADOTD = prhs[2], *ABLK = prhs[3], *BLK = prhs[4];
  MEX_REQUIRE(mxIsStruct(DENSE), "dense should be a structure.");
  sb_idx nl = (sb_idx)mxGetScalar(need_field(DENSE, "l", "Missing field dense.l."));
  const mxArray *Q = need_field(DENSE, "q", "Missing field dense.q.");
  sb_idx nq = (sb_idx)numel(Q);
  const mxArray *COLS = need_field(DENSE, "cols", "Missing field dense.cols.");
  sb_idx nden = (sb_idx)numel(COLS) - nl - nq;
  MEX_REQUIRE(nden >= 0, "dense.q size mismatch.");
  const mxArray *A = need_field(DENSE, "A", "Missing field dense.A.");
  MEX_REQUIRE(mxIsSparse(A), "dense.A must be sparse");
  mwSize m = mxGetM(A);
  MEX_REQUIRE((sb_idx)mxGetN(A) - nl == nq + nden, "dense.A size mismatch");
  MEX_REQUIRE(mxIsStruct(D), "d should be a structure.");
  const mxArray *Q1 = need_field(D, "q1", "Missing field d.q1."), *Q2 = need_field(D, "q2", "Missing field d.q2.");
  sb_idx lorN = (sb_idx)numel(Q1);
  MEX_REQUIRE(mxIsSparse(ADOTD), "sparAD must be sparse");
  MEX_REQUIRE((mxGetM(ADOTD) == m || nq <= 0) && (sb_idx)mxGetN(ADOTD) == nq, "Size mismatch sparAD");
  MEX_REQUIRE((sb_idx)numel(BLK) == lorN + 1, "blkstart size mismatch");
  plhs[0] = mxDuplicateArray(ABLK);                                         // Ad = Ablk
  if (nq == 0) return;
  std::vector<sb_idx> q, cols, bs;
  idx_from_double(Q, q, 1, "dense.q");
  idx_from_double(COLS, cols, 1, "dense.cols");
  idx_from_double(BLK, bs, 1, "blkstart");
  const sb_idx firstQ = bs[0];
  const double *d1 = mxGetPr(Q1), *d2 = mxGetPr(Q2);
  std::vector<double> d1q(nq), d2c(nden ? nden : 1);
  std::vector<sb_idx> colbeg(nq + 1, 0);
  sb_idx j = 0;
  for (sb_idx k = 0; k < nq; k++) {                                          // adendotd.c:96-118
    MEX_REQUIRE(q[k] < lorN, "dense.q out of range");
    d1q[k] = d1[q[k]];
    colbeg[k] = j;
    const sb_idx blkend = bs[q[k] + 1];
    while (j < nden && cols[nl + nq + j] < blkend) { d2c[j] = d2[cols[nl + nq + j] - firstQ]; j++; }
  }
  colbeg[nq] = j;
  const mwIndex *ajc = mxGetJc(A) + nl;
  std::vector<sb_idx> ajc_rel(nq + nden + 1);
  for (sb_idx c = 0; c <= nq + nden; c++) ajc_rel[c] = (sb_idx)(ajc[c] - ajc[0]);
  int rc = sb200_adendotd((sb_idx)m, nq, nden, as_idx(mxGetJc(plhs[0])), as_idx(mxGetIr(plhs[0])), as_idx(mxGetJc(ADOTD)),
                          as_idx(mxGetIr(ADOTD)), mxGetPr(ADOTD), ajc_rel.data(), as_idx(mxGetIr(A)) + ajc[0], mxGetPr(A) + ajc[0],
                          d1q.data(), colbeg.data(), d2c.data(), mxGetPr(plhs[0]));
  if (rc) { mxDestroyArray(plhs[0]); plhs[0] = NULL; sb_check(rc, "adendotd"); }
}
