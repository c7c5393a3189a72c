// y = qblkmul(mu,d,blkstart)   y[k] = mu(k)*d[k] per Lorentz block (qblkmul.c:44-49 signature, :57-116)
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 3, "qblkmul requires more input arguments.");
  MEX_REQUIRE(nlhs <= 1, "qblkmul generates 1 output argument.");
  const mxArray *MU = prhs[0], *D = prhs[1], *BLK = prhs[2];
  sb_idx nblk = (sb_idx)numel(MU);
  MEX_REQUIRE((sb_idx)numel(BLK) == nblk + 1, "blkstart size mismatch.");
  std::vector<sb_idx> bs;
  idx_from_double(BLK, bs, 1, "blkstart");
  const double *d = mxGetPr(D);
  sb_idx qDim = (sb_idx)numel(D), span = bs[nblk] - bs[0];
  if (qDim != span) {
    if (qDim == nblk + span) d += nblk;
    else { MEX_REQUIRE(qDim >= bs[nblk], "d size mismatch."); d += bs[0]; }
  }
  std::vector<sb_idx> rel(nblk + 1);
  for (sb_idx k = 0; k <= nblk; k++) rel[k] = bs[k] - bs[0];
  plhs[0] = mxCreateDoubleMatrix((mwSize)span, 1, mxREAL);
  int rc = sb200_qblkmul(nblk, rel.data(), mxGetPr(MU), d, mxGetPr(plhs[0]));
  if (rc) { mxDestroyArray(plhs[0]); plhs[0] = NULL; sb_check(rc, "qblkmul"); }
}
