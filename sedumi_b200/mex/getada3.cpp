// [ADA,absd] = getada3(ADA,At,Ajc1,Aord,udsqr,K)   PSD part of A*D(d^2)*A', absd, symmetrisation
// (getada3.c:50-60 signature, :370-569 mexFunction)
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 6, "getADA requires more input arguments.");
  MEX_REQUIRE(nlhs <= 2, "getADA produces less output arguments.");
  const mxArray *ADA = prhs[0], *AT = prhs[1], *AJC1 = prhs[2], *AORD = prhs[3], *UDSQR = prhs[4];
  ConeK K;
  read_cone(prhs[5], K);
  sb_idx lenud = K.rDim + K.hDim;                 // Hermitian blocks: [vec Re; vec Im] (getada3.c:404-412, spscale.c:472-480)
  sb_idx lenfull = K.lpN + K.qDim + lenud;
  const mxArray *bsf = need_field(prhs[5], "blkstart", "Missing K.blkstart.");
  MEX_REQUIRE(numel(bsf) == (mwSize)(2 + K.lorN + K.sdpN), "Size mismatch K.blkstart.");
  std::vector<sb_idx> bs_all, blkstart;
  idx_from_double(bsf, bs_all, 1, "K.blkstart");
  for (sb_idx k = 0; k <= K.sdpN; k++) blkstart.push_back(bs_all[K.lorN + 1 + k]);
  MEX_REQUIRE(mxGetM(AT) == (mwSize)lenfull, "Size mismatch At");
  mwSize m = mxGetN(AT);
  MEX_REQUIRE(mxIsSparse(AT), "At should be sparse.");
  MEX_REQUIRE(numel(UDSQR) == (mwSize)lenud, "udsqr size mismatch.");
  MEX_REQUIRE(numel(AJC1) == m, "Ajc1 size mismatch");
  MEX_REQUIRE(mxIsStruct(AORD), "Aord should be a structure.");
  const mxArray *DZ = need_field(AORD, "dz", "Missing field Aord.dz.");
  MEX_REQUIRE(mxGetN(DZ) >= m, "Size mismatch Aord.dz.");
  MEX_REQUIRE(mxIsSparse(DZ), "Aord.dz should be sparse.");
  const mxArray *SP = need_field(AORD, "sperm", "Missing field Aord.sperm.");
  MEX_REQUIRE(numel(SP) == m, "Aord.sperm size mismatch");
  MEX_REQUIRE(mxGetM(ADA) == m && mxGetN(ADA) == m, "Size mismatch ADA.");
  MEX_REQUIRE(mxIsSparse(ADA), "ADA should be sparse.");
  std::vector<sb_idx> Ajc1, perm;
  idx_from_double(AJC1, Ajc1, 0, "Ajc1");
  idx_from_double(SP, perm, 1, "Aord.sperm");
  // number of leading constraints (in sperm order) without PSD nonzeros: their absd stays 0
  // (getada3.c:282-284); all of them if dz is empty (:273-274)
  const mwIndex *dzjc = mxGetJc(DZ);
  sb_idx first = (sb_idx)m;
  if (K.sdpN > 0 && dzjc[m] > 0) { first = 0; while (dzjc[first + 1] == 0) first++; }
  const mwIndex *adajc = mxGetJc(ADA), *adair = mxGetIr(ADA);
  // Lorentz layout is irrelevant here: describe only the PSD part to the plan
  std::vector<sb_idx> Ajc_lq(m);            // columns restricted to the PSD part: start = Ajc1
  sb200_ada_plan *pl = NULL;
  sb_check(sb200_ada_plan_get_h(&pl, (sb_idx)lenfull, (sb_idx)m, as_idx(mxGetJc(AT)), as_idx(mxGetIr(AT)), Ajc1.data(),
                                0, 0, NULL, K.sdpN, K.rsdpN, blkstart.data(), K.s.data(), as_idx(adajc), as_idx(adair)), "getada3");
  mxArray *out0 = sparse_with_pattern(m, m, adajc, adair);
  mxArray *out1 = mxCreateDoubleMatrix(m, 1, mxREAL);
  int rc = sb200_getada3(pl, mxGetPr(AT), mxGetPr(UDSQR), lenud, perm.data(), first, mxGetPr(ADA), mxGetPr(out0), mxGetPr(out1));
  if (rc) { mxDestroyArray(out0); mxDestroyArray(out1); sb_check(rc, "getada3"); }
  plhs[0] = out0;
  if (nlhs >= 2) plhs[1] = out1; else mxDestroyArray(out1);
}
