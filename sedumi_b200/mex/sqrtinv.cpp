// y = sqrtinv(q,vlab,K)   Y_k = (Q_k / diag(sqrt(vlab_k)))'   (sqrtinv.c:86-148)
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 3, "sqrtinv requires more input arguments.");
  MEX_REQUIRE(nlhs <= 1, "sqrtinv generates less output arguments.");
  ConeK K;
  read_cone(prhs[2], K);
  const sb_idx lenud = K.rDim + K.hDim, diagskip = K.lpN + 2 * K.lorN, lendiag = diagskip + K.rLen + K.hLen;
  MEX_REQUIRE((sb_idx)numel(prhs[1]) == lendiag, "v size mismatch");
  MEX_REQUIRE((sb_idx)numel(prhs[0]) == lenud, "q size mismatch");
  plhs[0] = mxCreateDoubleMatrix((mwSize)lenud, 1, mxREAL);
  int rc = sb200_sqrtinv(K.sdpN, K.rsdpN, K.s.data(), mxGetPr(prhs[0]), mxGetPr(prhs[1]) + diagskip, mxGetPr(plhs[0]));
  if (rc) { mxDestroyArray(plhs[0]); plhs[0] = NULL; sb_check(rc, "sqrtinv"); }
}
