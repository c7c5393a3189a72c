// [u,perm,gjc,g] = urotorder(u,K,maxu[,permIN])   stable re-ordering of the triangular factor
// (urotorder.c:46-57 signature, :312-490 mexFunction)
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 3, "urotorder requires more input arguments.");
  MEX_REQUIRE(nlhs <= 4, "urotorder generates less output arguments.");
  ConeK K;
  read_cone(prhs[1], K);
  const bool herm = K.rsdpN != K.sdpN;             // Hermitian blocks (urotorder.c:420-455)
  sb_idx lenud = K.rDim + K.hDim, sdplen = K.rLen + K.hLen;
  double maxu = mxGetScalar(prhs[2]);
  MEX_REQUIRE((sb_idx)numel(prhs[0]) == lenud, "u size mismatch");
  const double *permOld = NULL;
  if (nrhs >= 4 && numel(prhs[3]) > 0) {
    MEX_REQUIRE((sb_idx)numel(prhs[3]) == sdplen, "perm size mismatch");
    permOld = mxGetPr(prhs[3]);
  }
  sb_idx gworst = 0;
  for (sb_idx k = 0; k < K.sdpN; k++) gworst += k < K.rsdpN ? K.s[k] * (K.s[k] - 1) : 3 * K.s[k] * (K.s[k] - 1) / 2;
  mxArray *out[4];
  out[0] = mxCreateDoubleMatrix((mwSize)lenud, 1, mxREAL);
  out[1] = mxCreateDoubleMatrix((mwSize)sdplen, 1, mxREAL);
  out[2] = mxCreateDoubleMatrix((mwSize)sdplen, 1, mxREAL);
  std::vector<sb_idx> perm((size_t)(sdplen ? sdplen : 1)), gjc((size_t)(sdplen ? sdplen : 1));
  double *gw = (double *)mxCalloc((size_t)(gworst ? gworst : 1), sizeof(double));
  int rc = herm ? sb200_urotorder_h(K.sdpN, K.rsdpN, K.s.data(), mxGetPr(prhs[0]), maxu, mxGetPr(out[0]), perm.data(), gjc.data(), gw)
                : sb200_urotorder(K.sdpN, K.s.data(), mxGetPr(prhs[0]), maxu, mxGetPr(out[0]), perm.data(), gjc.data(), gw);
  if (rc) { mxFree(gw); for (int i = 0; i < 3; i++) mxDestroyArray(out[i]); sb_check(rc, "urotorder"); }
  double *permPr = mxGetPr(out[1]), *gjcPr = mxGetPr(out[2]);
  sb_idx inz = 0, poff = 0, goff = 0;
  for (sb_idx k = 0; k < K.sdpN; k++) {                 // pack g; compose perm (urotorder.c:407-417)
    sb_idx nk = K.s[k];
    for (sb_idx i = 0; i < nk; i++) {
      permPr[poff + i] = permOld ? permOld[poff + perm[poff + i]] : 1.0 + (double)perm[poff + i];
      gjcPr[poff + i] = (double)gjc[poff + i];
    }
    const bool cplx = k >= K.rsdpN;                     // rotations: 2 doubles, or 3 for a Hermitian block
    sb_idx cnt = (cplx ? 3 : 2) * gjc[poff + nk - 1];
    memmove(gw + inz, gw + goff, (size_t)cnt * sizeof(double));
    inz += cnt; poff += nk; goff += cplx ? 3 * nk * (nk - 1) / 2 : nk * (nk - 1);
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
