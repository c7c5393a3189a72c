// smult = adenscale(dense,d,blkstart)   smult(j) = det(d_k) for every dense norm-bound column j of
// Lorentz block k (adenscale.c:41-47 signature, :85-163 mexFunction; called from deninfac.m:62).
// nden values of pure index bookkeeping: no arithmetic, nothing to launch.
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 3, "adenscale requires more input arguments");
  MEX_REQUIRE(nlhs <= 1, "adenscale produces less output arguments");
  const mxArray *DENSE = prhs[0], *D = prhs[1], *BLK = prhs[2];
  MEX_REQUIRE(mxIsStruct(DENSE), "dense should be a structure.");
  sb_idx nl = (sb_idx)mxGetScalar(need_field(DENSE, "l", "Missing field dense.l."));
  const mxArray *Q = need_field(DENSE, "q", "Missing field dense.q.");
  sb_idx nq = (sb_idx)numel(Q);
  const mxArray *COLS = need_field(DENSE, "cols", "Missing field dense.cols.");
  sb_idx nden = (sb_idx)numel(COLS) - (nq + nl);
  MEX_REQUIRE(nden >= 0, "dense.cols size mismatch.");
  MEX_REQUIRE(mxIsStruct(D), "d should be a structure.");
  const mxArray *DET = need_field(D, "det", "Missing field d.det.");
  sb_idx lorN = (sb_idx)numel(DET);
  MEX_REQUIRE((sb_idx)numel(BLK) == lorN + 1, "blkstart size mismatch");
  std::vector<sb_idx> q, cols, bs;
  idx_from_double(Q, q, 1, "dense.q");
  idx_from_double(COLS, cols, 1, "dense.cols");
  idx_from_double(BLK, bs, 1, "blkstart");
  plhs[0] = mxCreateDoubleMatrix((mwSize)nden, 1, mxREAL);
  double *smult = mxGetPr(plhs[0]);
  const double *detd = mxGetPr(DET);
  sb_idx j = 0;
  for (sb_idx k = 0; k < nq; k++) {
    MEX_REQUIRE(q[k] < lorN, "dense.q out of range");
    const double detdk = detd[q[k]];
    const sb_idx blkend = bs[q[k] + 1];
    while (j < nden && cols[nq + nl + j] < blkend) smult[j++] = detdk;
  }
}
