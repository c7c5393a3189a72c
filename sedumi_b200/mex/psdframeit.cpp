// x = psdframeit(lab,frms,K)   X_k = Qb'*diag(lab_k)*Qb, Qb in Householder product form
// (psdframeit.c:43-49 signature, :106-168 mexFunction)
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 3, "psdframeit requires more input arguments.");
  MEX_REQUIRE(nlhs <= 1, "psdframeit generates less output arguments.");
  ConeK K;
  read_cone(prhs[2], K);
  if (K.rsdpN != K.sdpN) mexErrMsgTxt("psdframeit: Hermitian PSD blocks are not supported by the B200 plugin yet.");
  sb_idx lenud = K.rDim, lendiag = K.lpN + 2 * K.lorN + K.rLen;
  const double *lab = mxGetPr(prhs[0]);
  if ((sb_idx)numel(prhs[0]) != K.rLen) {
    MEX_REQUIRE((sb_idx)numel(prhs[0]) == lendiag, "lab size mismatch");
    lab += K.lpN + 2 * K.lorN;
  }
  MEX_REQUIRE((sb_idx)numel(prhs[1]) == lenud, "frms size mismatch");
  plhs[0] = mxCreateDoubleMatrix((mwSize)lenud, 1, mxREAL);
  int rc = sb200_psdframeit(K.sdpN, K.s.data(), lab, mxGetPr(prhs[1]), mxGetPr(plhs[0]));
  if (rc) { mxDestroyArray(plhs[0]); plhs[0] = NULL; sb_check(rc, "psdframeit"); }
}
