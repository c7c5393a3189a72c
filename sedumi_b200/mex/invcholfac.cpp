// y = invcholfac(u,K[,perm])   per PSD block: Y(perm,perm) = triu(U)'*triu(U)
// (invcholfac.c:42-51 signature, :59-168 mexFunction)
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 2, "invcholfac requires at least 2 input arguments.");
  MEX_REQUIRE(nlhs <= 1, "invcholfac generates 1 output argument.");
  ConeK K;
  read_cone(prhs[1], K);
  const bool herm = K.rsdpN != K.sdpN;             // Hermitian blocks: [vec Re; vec Im] each (invcholfac.c:131-141)
  sb_idx lenud = K.rDim + K.hDim;
  MEX_REQUIRE(numel(prhs[0]) == (mwSize)lenud, "u size mismatch");
  bool isperm = nrhs >= 3 && numel(prhs[2]) > 0;
  std::vector<sb_idx> perm;
  if (isperm) {
    MEX_REQUIRE(numel(prhs[2]) >= (mwSize)(K.rLen + K.hLen), "perm size mismatch");
    idx_from_double(prhs[2], perm, 1, "perm");
  }
  plhs[0] = mxCreateDoubleMatrix((mwSize)lenud, 1, mxREAL);
  int rc = herm ? sb200_invcholfac_h(K.sdpN, K.rsdpN, K.s.data(), mxGetPr(prhs[0]), isperm ? perm.data() : NULL, mxGetPr(plhs[0]))
                : sb200_invcholfac(K.sdpN, K.s.data(), mxGetPr(prhs[0]), isperm ? perm.data() : NULL, mxGetPr(plhs[0]));
  if (rc) { mxDestroyArray(plhs[0]); plhs[0] = NULL; sb_check(rc, "invcholfac"); }
}
