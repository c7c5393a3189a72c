// y = psdscale(ud,x,K[,transp])   per PSD block Y = T'*X*T, T = tril(U) or triu(U)
// The reference implements this in M (psdscale.m:45-119); a MEX of the same name in the
// same directory takes precedence, so this plugin is a drop-in for the .m file.
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 3, "psdscale requires at least 3 input arguments.");
  MEX_REQUIRE(nlhs <= 1, "psdscale generates 1 output argument.");
  ConeK K;
  read_cone(prhs[2], K);
  if (K.sdpN == 0) { plhs[0] = mxCreateDoubleMatrix(0, 0, mxREAL); return; }      // y = [] (psdscale.m:47-50)
  const bool herm = K.rsdpN != K.sdpN;             // Hermitian blocks: [vec Re; vec Im] each (psdscale.m:55-58,68-72)
  bool transp = nrhs >= 4 && numel(prhs[3]) > 0 && mxGetScalar(prhs[3]) != 0.0;
  const mxArray *UD = prhs[0], *ufield = UD, *pfield = NULL;
  if (mxIsStruct(UD)) {
    ufield = need_field(UD, "u", "Missing field ud.u.");
    pfield = mxGetField(UD, 0, "perm");
  }
  sb_idx N = K.rDim + K.hDim;
  MEX_REQUIRE(numel(ufield) >= (mwSize)N, "ud.u size mismatch");
  MEX_REQUIRE(numel(prhs[1]) >= (mwSize)N, "x size mismatch");
  std::vector<sb_idx> perm;
  bool isperm = pfield && numel(pfield) > 0;
  if (isperm) {
    MEX_REQUIRE(numel(pfield) >= (mwSize)(K.rLen + K.hLen), "ud.perm size mismatch");
    idx_from_double(pfield, perm, 1, "ud.perm");
  }
  // PSD part is the tail (psdscale.m:58).  psdscale.m takes any x, sparse included (it even sparsifies blocks
  // itself, :100-102): a sparse vector is scattered into a dense tail here.
  const mwSize tail0 = numel(prhs[1]) - (mwSize)N;
  std::vector<double> xdense;
  const double *x;
  if (mxIsSparse(prhs[1])) {
    xdense.assign((size_t)N, 0.0);
    const mwIndex *xjc = mxGetJc(prhs[1]), *xir = mxGetIr(prhs[1]);
    const double *xpr = mxGetPr(prhs[1]);
    const mwSize xm = mxGetM(prhs[1]), xn = mxGetN(prhs[1]);
    for (mwSize c = 0; c < xn; c++)
      for (mwIndex p = xjc[c]; p < xjc[c + 1]; p++) {
        const mwSize lin = c * xm + xir[p];
        if (lin >= tail0) xdense[lin - tail0] = xpr[p];
      }
    x = xdense.data();
  } else x = mxGetPr(prhs[1]) + tail0;
  plhs[0] = mxCreateDoubleMatrix((mwSize)N, 1, mxREAL);
  int rc = herm ? sb200_psdscale_h(K.sdpN, K.rsdpN, K.s.data(), mxGetPr(ufield), isperm ? perm.data() : NULL, x, transp ? 1 : 0, mxGetPr(plhs[0]))
                : sb200_psdscale(K.sdpN, K.s.data(), mxGetPr(ufield), isperm ? perm.data() : NULL, x, transp ? 1 : 0, mxGetPr(plhs[0]));
  if (rc) { mxDestroyArray(plhs[0]); plhs[0] = NULL; sb_check(rc, "psdscale"); }
}
