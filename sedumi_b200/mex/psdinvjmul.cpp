// z = psdinvjmul(xlab,xfrm,y,K)   solves X*Z + Z*X = 2*Y per PSD block, X = Qb'*diag(xlab)*Qb
// (psdinvjmul.c:46-54 signature, :166-227 mexFunction)
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 4, "psdinvjmul requires more input arguments.");
  MEX_REQUIRE(nlhs <= 1, "psdinvjmul generates 1 output argument.");
  ConeK K;
  read_cone(prhs[3], K);
  const bool herm = K.rsdpN != K.sdpN;             // Hermitian blocks (psdinvjmul.c:131-155,186-199)
  sb_idx lenud = K.rDim + K.hDim, lenfull = K.lpN + K.qDim + lenud, lendiag = K.lpN + 2 * K.lorN + K.rLen + K.hLen;
  MEX_REQUIRE(!mxIsSparse(prhs[0]) && !mxIsSparse(prhs[2]), "Sparse inputs not supported by this version of psdinvjmul.");
  const double *x = mxGetPr(prhs[0]), *y = mxGetPr(prhs[2]);
  if ((sb_idx)numel(prhs[2]) != lenud) { MEX_REQUIRE((sb_idx)numel(prhs[2]) == lenfull, "size y mismatch."); y += K.lpN + K.qDim; }
  if ((sb_idx)numel(prhs[0]) != K.rLen + K.hLen) { MEX_REQUIRE((sb_idx)numel(prhs[0]) == lendiag, "size xlab mismatch."); x += K.lpN + 2 * K.lorN; }
  MEX_REQUIRE((sb_idx)numel(prhs[1]) == lenud + K.hLen, "size xfrm mismatch.");
  plhs[0] = mxCreateDoubleMatrix((mwSize)lenud, 1, mxREAL);
  int rc = herm ? sb200_psdinvjmul_h(K.sdpN, K.rsdpN, K.s.data(), x, mxGetPr(prhs[1]), y, mxGetPr(plhs[0]))
                : sb200_psdinvjmul(K.sdpN, K.s.data(), x, mxGetPr(prhs[1]), y, mxGetPr(plhs[0]));
  if (rc) { mxDestroyArray(plhs[0]); plhs[0] = NULL; sb_check(rc, "psdinvjmul"); }
}
