// x = psdframeit(lab,frms,K)   X_k = Qb'*diag(lab_k)*Qb, Qb in Householder product form
// (psdframeit.c:43-49 signature, :106-168 mexFunction)
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 3, "psdframeit requires more input arguments.");
  MEX_REQUIRE(nlhs <= 1, "psdframeit generates less output arguments.");
  ConeK K;
  read_cone(prhs[2], K);
  const bool herm = K.rsdpN != K.sdpN;             // Hermitian blocks: frames [Re c | Im c | beta] (psdframeit.c:80-97,128)
  sb_idx lenud = K.rDim + K.hDim, lendiag = K.lpN + 2 * K.lorN + K.rLen + K.hLen;
  const double *lab = mxGetPr(prhs[0]);
  if ((sb_idx)numel(prhs[0]) != K.rLen + K.hLen) {
    MEX_REQUIRE((sb_idx)numel(prhs[0]) == lendiag, "lab size mismatch");
    lab += K.lpN + 2 * K.lorN;
  }
  MEX_REQUIRE((sb_idx)numel(prhs[1]) == lenud + K.hLen, "frms size mismatch");
  plhs[0] = mxCreateDoubleMatrix((mwSize)lenud, 1, mxREAL);
  int rc = herm ? sb200_psdframeit_h(K.sdpN, K.rsdpN, K.s.data(), lab, mxGetPr(prhs[1]), mxGetPr(plhs[0]))
                : sb200_psdframeit(K.sdpN, K.s.data(), lab, mxGetPr(prhs[1]), mxGetPr(plhs[0]));
  if (rc) { mxDestroyArray(plhs[0]); plhs[0] = NULL; sb_check(rc, "psdframeit"); }
}
