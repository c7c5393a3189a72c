// ADA = getada2(ADA,DAt,Aord,K)   ADA += DAt.q'*DAt.q   (getada2.c:48-55 signature, :126-214)
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 4, "getADA requires more input arguments.");
  MEX_REQUIRE(nlhs <= 1, "getADA produces less output arguments.");
  const mxArray *ADA = prhs[0], *DAT = prhs[1], *AORD = prhs[2];
  ConeK K;
  read_cone(prhs[3], K);
  mwSize m = mxGetM(ADA);
  MEX_REQUIRE(mxGetN(ADA) == m, "Size mismatch ADA.");
  MEX_REQUIRE(mxIsSparse(ADA), "ADA should be sparse.");
  plhs[0] = sparse_with_pattern(m, m, mxGetJc(ADA), mxGetIr(ADA));
  // ADA_OUT starts as a duplicate of ADA_IN (getada2.c:153 mxDuplicateArray): every early exit below returns the
  // input values unchanged -- sparse_with_pattern leaves the value array uninitialised on purpose
  memcpy(mxGetPr(plhs[0]), mxGetPr(ADA), mxGetJc(ADA)[m] * sizeof(double));
  if (K.lorN <= 0) return;                                  // ready if no Lorentz blocks (getada2.c:151-152)
  MEX_REQUIRE(mxIsStruct(DAT), "DAt should be a structure.");
  const mxArray *Q = need_field(DAT, "q", "Missing field DAt.q.");
  MEX_REQUIRE(mxGetM(Q) == (mwSize)K.lorN && mxGetN(Q) == m, "Size mismatch DAt.q");
  MEX_REQUIRE(mxIsSparse(Q), "DAt.q should be sparse.");
  MEX_REQUIRE(mxIsStruct(AORD), "Aord should be a structure.");
  const mxArray *QP = need_field(AORD, "qperm", "Missing field Aord.qperm.");
  MEX_REQUIRE(numel(QP) == m, "Size mismatch Aord.qperm.");
  if (mxGetJc(Q)[m] == 0) return;                           // DAt.q empty (all Lorentz cones dense, getDAtm.m:44): ADA unchanged
  std::vector<sb_idx> perm;
  idx_from_double(QP, perm, 1, "Aord.qperm");
  const mwIndex *adajc = mxGetJc(ADA), *adair = mxGetIr(ADA);
  std::vector<sb_idx> zer(m + 1, 0);
  sb200_ada_plan *pl = NULL;
  int rc = sb200_ada_plan_get(&pl, 0, (sb_idx)m, zer.data(), zer.data(), zer.data(), 0, 0, NULL, 0, NULL, NULL,
                              as_idx(adajc), as_idx(adair));
  if (!rc) rc = sb200_getada2(pl, K.lorN, as_idx(mxGetJc(Q)), as_idx(mxGetIr(Q)), mxGetPr(Q), perm.data(),
                              mxGetPr(ADA), mxGetPr(plhs[0]));
  if (rc) { mxDestroyArray(plhs[0]); plhs[0] = NULL; sb_check(rc, "getada2"); }
}
