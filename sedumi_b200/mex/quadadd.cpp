// [zhi,zlo] = quadadd(xhi,xlo,y)   double-double accumulate (quadadd.c:45-52 signature, :89-130)
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 3, "quadadd requires more input arguments.");
  MEX_REQUIRE(nlhs <= 2, "quadadd produces less output arguments.");
  mwSize m = mxGetM(prhs[0]);
  MEX_REQUIRE(mxGetM(prhs[1]) == m, "xlo size mismatch.");
  MEX_REQUIRE(mxGetM(prhs[2]) == m, "y size mismatch.");
  mxArray *zhi = mxCreateDoubleMatrix(m, 1, mxREAL), *zlo = mxCreateDoubleMatrix(m, 1, mxREAL);
  int rc = sb200_quadadd((sb_idx)m, mxGetPr(prhs[0]), mxGetPr(prhs[1]), mxGetPr(prhs[2]), mxGetPr(zhi), mxGetPr(zlo));
  if (rc) { mxDestroyArray(zhi); mxDestroyArray(zlo); sb_check(rc, "quadadd"); }
  plhs[0] = zhi;
  if (nlhs >= 2) plhs[1] = zlo; else mxDestroyArray(zlo);
}
