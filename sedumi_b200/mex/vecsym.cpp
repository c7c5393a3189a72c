// y = vecsym(x,K)   Y_k = (X_k + X_k')/2 on the PSD blocks, LP / Lorentz part copied   (vecsym.c:139-175)
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 2, "vecsym requires 2 input arguments.");
  MEX_REQUIRE(nlhs <= 1, "vecsym generates 1 output argument.");
  ConeK K;
  read_cone(prhs[1], K);
  const sb_idx lq = K.lpN + K.qDim, lenfull = lq + K.rDim + K.hDim;
  MEX_REQUIRE(!mxIsSparse(prhs[0]), "x must be full.");
  MEX_REQUIRE((sb_idx)numel(prhs[0]) == lenfull, "Parameter `x' size mismatch.");
  plhs[0] = mxCreateDoubleMatrix((mwSize)lenfull, 1, mxREAL);
  const double *x = mxGetPr(prhs[0]);
  double *y = mxGetPr(plhs[0]);
  memcpy(y, x, lq * sizeof(double));
  int rc = sb200_vecsym(K.sdpN, K.rsdpN, K.s.data(), x + lq, y + lq);
  if (rc) { mxDestroyArray(plhs[0]); plhs[0] = NULL; sb_check(rc, "vecsym"); }
}
