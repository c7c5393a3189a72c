// y = bwdpr1(Lden,b)   product-form backward solve over the dense-column factors
// (bwdpr1.c signature: bwdpr1.c:47-53, mexFunction :171-275)
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 2, "bwdpr1 requires more input arguments.");
  MEX_REQUIRE(nlhs <= 1, "bwdpr1 generates less output arguments.");
  const mxArray *LDEN = prhs[0], *B = prhs[1];
  MEX_REQUIRE(!mxIsSparse(B), "b should be full");
  mwSize m = mxGetM(B), n = mxGetN(B);
  plhs[0] = mxDuplicateArray(B);
  MEX_REQUIRE(mxIsStruct(LDEN), "Parameter `Lden' should be a structure.");
  const mxArray *BJ = need_field(LDEN, "betajc", "Missing field Lden.betajc.");
  if (numel(BJ) <= 1) return;                         // no dense columns: y = b (bwdpr1.c:206-208)
  mwSize nden = numel(BJ) - 1;
  const mxArray *P = need_field(LDEN, "p", "Missing field Lden.p.");
  const mxArray *DP = need_field(LDEN, "dopiv", "Missing field Lden.dopiv.");
  MEX_REQUIRE(numel(DP) == nden, "Size mismatch Lden.dopiv.");
  const mxArray *PP = need_field(LDEN, "pivperm", "Missing field Lden.pivperm.");
  const mxArray *BE = need_field(LDEN, "beta", "Missing field Lden.beta.");
  const mxArray *DZ = need_field(LDEN, "dz", "Missing field Lden.dz.");
  MEX_REQUIRE(mxGetM(DZ) == m && mxGetN(DZ) == nden, "Lden.dz size mismatch.");
  MEX_REQUIRE(mxIsSparse(DZ), "Lden.dz must be sparse.");
  std::vector<sb_idx> betajc, pivperm;
  idx_from_double(BJ, betajc, 1, "Lden.betajc");
  idx_from_double(PP, pivperm, 0, "Lden.pivperm");
  MEX_REQUIRE((sb_idx)numel(BE) == betajc[nden], "Size mismatch Lden.beta.");
  int rc = sb200_dpr1solve(1, (sb_idx)m, (sb_idx)n, (sb_idx)nden, as_idx(mxGetJc(DZ)), as_idx(mxGetIr(DZ)), mxGetPr(P),
                           pivperm.data(), (sb_idx)pivperm.size(), mxGetPr(BE), betajc.data(), mxGetPr(DP), mxGetPr(plhs[0]));
  if (rc) { mxDestroyArray(plhs[0]); plhs[0] = NULL; sb_check(rc, "bwdpr1"); }
}
