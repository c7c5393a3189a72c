// ADA = getada1(ADA,At,Ajc2,perm,d,blkstart)   LP + Lorentz-det part of A*D(d^2)*A'
// (getada1.c:54-63 signature, :161-261 mexFunction)
#include <time.h>
#include <stdlib.h>
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 6, "getADA requires more input arguments.");
  MEX_REQUIRE(nlhs <= 1, "getADA produces less output arguments.");
  const mxArray *ADA = prhs[0], *AT = prhs[1], *AJC2 = prhs[2], *PERM = prhs[3], *D = prhs[4], *BLK = prhs[5];
  mwSize nblk = numel(BLK);
  MEX_REQUIRE(nblk >= 1, "Size mismatch blkstart.");
  mwSize m = mxGetN(AT);
  MEX_REQUIRE(mxIsSparse(AT), "At should be sparse.");
  MEX_REQUIRE(numel(AJC2) == m, "Size mismatch Ajc2.");
  MEX_REQUIRE(numel(PERM) == m, "Size mismatch perm.");
  MEX_REQUIRE(mxIsStruct(D), "Parameter `d' should be a structure.");
  const mxArray *dl = need_field(D, "l", "Field d.l missing.");
  const mxArray *ddet = need_field(D, "det", "Field d.det missing.");
  std::vector<sb_idx> qstart, Ajc2, perm;
  idx_from_double(BLK, qstart, 1, "blkstart");
  idx_from_double(AJC2, Ajc2, 0, "Ajc2");
  idx_from_double(PERM, perm, 1, "perm");
  sb_idx lpN = (sb_idx)numel(dl), nq = (sb_idx)nblk - 1;
  MEX_REQUIRE(numel(ddet) == (mwSize)nq, "Size d.det mismatch");
  MEX_REQUIRE(qstart[0] == lpN + nq, "blkstart mismatches d.l / d.det");
  MEX_REQUIRE(mxGetM(AT) >= (mwSize)qstart[nq], "Size mismatch At");
  MEX_REQUIRE(mxGetM(ADA) == m && mxGetN(ADA) == m, "Size mismatch ADA.");
  MEX_REQUIRE(mxIsSparse(ADA), "ADA should be sparse.");
  const mwIndex *adajc = mxGetJc(ADA), *adair = mxGetIr(ADA);
  const bool trace = getenv("SB200_TRACE") != NULL;
  struct timespec ts0, ts1, ts2;
  clock_gettime(CLOCK_MONOTONIC, &ts0);
  sb200_ada_plan *pl = NULL;
  sb_check(sb200_ada_plan_get(&pl, (sb_idx)mxGetM(AT), (sb_idx)m, as_idx(mxGetJc(AT)), as_idx(mxGetIr(AT)), Ajc2.data(),
                              lpN, nq, qstart.data(), 0, NULL, NULL, as_idx(adajc), as_idx(adair)), "getada1");
  clock_gettime(CLOCK_MONOTONIC, &ts1);
  plhs[0] = sparse_with_pattern(m, m, adajc, adair);
  clock_gettime(CLOCK_MONOTONIC, &ts2);
  if (trace) fprintf(stderr, "[getada1 stub] plan lookup %.3f ms, output array %.3f ms\n", (ts1.tv_sec - ts0.tv_sec) * 1e3 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-6,
                     (ts2.tv_sec - ts1.tv_sec) * 1e3 + (ts2.tv_nsec - ts1.tv_nsec) * 1e-6);
  int rc = sb200_getada1(pl, mxGetPr(AT), perm.data(), mxGetPr(dl), mxGetPr(ddet), mxGetPr(plhs[0]));
  if (rc) { mxDestroyArray(plhs[0]); plhs[0] = NULL; sb_check(rc, "getada1"); }
}
