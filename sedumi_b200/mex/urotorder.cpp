// [u,perm,gjc,g] = urotorder(u,K,maxu[,permIN])   stable re-ordering of the triangular factor
// (urotorder.c:46-57 signature, :312-490 mexFunction)
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 3, "urotorder requires more input arguments.");
  MEX_REQUIRE(nlhs <= 4, "urotorder generates less output arguments.");
  ConeK K;
  read_cone(prhs[1], K);
  if (K.rsdpN != K.sdpN) mexErrMsgTxt("urotorder: Hermitian PSD blocks are not supported by the B200 plugin yet.");
  sb_idx lenud = K.rDim, sdplen = K.rLen;
  double maxu = mxGetScalar(prhs[2]);
  MEX_REQUIRE((sb_idx)numel(prhs[0]) == lenud, "u size mismatch");
  const double *permOld = NULL;
  if (nrhs >= 4 && numel(prhs[3]) > 0) {
    MEX_REQUIRE((sb_idx)numel(prhs[3]) == sdplen, "perm size mismatch");
    permOld = mxGetPr(prhs[3]);
  }
  sb_idx gworst = 0;
  for (sb_idx k = 0; k < K.sdpN; k++) gworst += K.s[k] * (K.s[k] - 1);
  mxArray *out[4];
  out[0] = mxCreateDoubleMatrix((mwSize)lenud, 1, mxREAL);
  out[1] = mxCreateDoubleMatrix((mwSize)sdplen, 1, mxREAL);
  out[2] = mxCreateDoubleMatrix((mwSize)sdplen, 1, mxREAL);
  std::vector<sb_idx> perm((size_t)(sdplen ? sdplen : 1)), gjc((size_t)(sdplen ? sdplen : 1));
  double *gw = (double *)mxCalloc((size_t)(gworst ? gworst : 1), sizeof(double));
  int rc = sb200_urotorder(K.sdpN, K.s.data(), mxGetPr(prhs[0]), maxu, mxGetPr(out[0]), perm.data(), gjc.data(), gw);
  if (rc) { mxFree(gw); for (int i = 0; i < 3; i++) mxDestroyArray(out[i]); sb_check(rc, "urotorder"); }
  double *permPr = mxGetPr(out[1]), *gjcPr = mxGetPr(out[2]);
  sb_idx inz = 0, poff = 0, goff = 0;
  for (sb_idx k = 0; k < K.sdpN; k++) {                 // pack g; compose perm (urotorder.c:407-417)
    sb_idx nk = K.s[k];
    for (sb_idx i = 0; i < nk; i++) {
      permPr[poff + i] = permOld ? permOld[poff + perm[poff + i]] : 1.0 + (double)perm[poff + i];
      gjcPr[poff + i] = (double)gjc[poff + i];
    }
    sb_idx cnt = 2 * gjc[poff + nk - 1];
    memmove(gw + inz, gw + goff, (size_t)cnt * sizeof(double));
    inz += cnt; poff += nk; goff += nk * (nk - 1);
  }
  // g: length-inz column vector whose buffer comes from the MEX allocator (urotorder.c:459-475)
  out[3] = mxCreateDoubleMatrix(1, 1, mxREAL);
  mxFree(mxGetPr(out[3]));
  double *gfin = NULL;
  if (inz > 0) { gfin = (double *)mxRealloc(gw, (size_t)inz * sizeof(double)); if (!gfin) mexErrMsgTxt("Memory allocation error."); }
  else mxFree(gw);
  mxSetPr(out[3], gfin);
  mxSetM(out[3], (mwSize)inz);
  int nout = nlhs > 1 ? nlhs : 1;
  for (int i = 0; i < 4; i++) { if (i < nout) plhs[i] = out[i]; else mxDestroyArray(out[i]); }
}
