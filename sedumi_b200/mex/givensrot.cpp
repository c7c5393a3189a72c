// y = givensrot(gjc,g,x,K)   apply the rotation list to every column of each PSD block
// (givensrot.c:44-51 signature, :93-168 mexFunction)
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 4, "givensrot requires more input arguments.");
  MEX_REQUIRE(nlhs <= 1, "givensrot generates less output arguments.");
  ConeK K;
  read_cone(prhs[3], K);
  const bool herm = K.rsdpN != K.sdpN;             // Hermitian blocks (givensrot.c:154-163)
  sb_idx lenud = K.rDim + K.hDim, sdplen = K.rLen + K.hLen;
  MEX_REQUIRE((sb_idx)mxGetM(prhs[2]) == lenud && (lenud == 0 || mxGetN(prhs[2]) == 1), "x size mismatch");
  MEX_REQUIRE((sb_idx)numel(prhs[0]) >= sdplen, "gjc size mismatch");
  std::vector<sb_idx> gjc;
  idx_from_double(prhs[0], gjc, 0, "gjc");
  plhs[0] = mxCreateDoubleMatrix((mwSize)lenud, 1, mxREAL);
  int rc = herm ? sb200_givensrot_h(K.sdpN, K.rsdpN, K.s.data(), gjc.data(), mxGetPr(prhs[1]), (sb_idx)numel(prhs[1]), mxGetPr(prhs[2]), mxGetPr(plhs[0]))
                : sb200_givensrot(K.sdpN, K.s.data(), gjc.data(), mxGetPr(prhs[1]), (sb_idx)numel(prhs[1]), mxGetPr(prhs[2]), mxGetPr(plhs[0]));
  if (rc) { mxDestroyArray(plhs[0]); plhs[0] = NULL; sb_check(rc, "givensrot"); }
}
