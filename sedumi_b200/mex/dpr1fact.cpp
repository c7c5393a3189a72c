// [Lden,d] = dpr1fact(x,d,Lsymb,smult,maxu)   product-form factor of diag(d) + sum smult_k p_k p_k'
// (dpr1fact.c:54-63 signature, :630-848 mexFunction)
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 5, "dpr1fact requires more input arguments");
  MEX_REQUIRE(nlhs <= 2, "dpr1fact produces less output arguments");
  const mxArray *X = prhs[0], *D = prhs[1], *LS = prhs[2], *SM = prhs[3];
  mwSize m = mxGetM(X), n = mxGetN(X);
  MEX_REQUIRE(mxIsSparse(X), "x should be sparse.");
  MEX_REQUIRE(numel(D) == m, "Size mismatch d.");
  MEX_REQUIRE(numel(SM) == n, "Size mismatch smult.");
  double maxu = mxGetScalar(prhs[4]);
  MEX_REQUIRE(mxIsStruct(LS), "Lsymb should be a structure.");
  const mxArray *DZ = need_field(LS, "dz", "Missing field Lsymb.dz.");
  MEX_REQUIRE(mxGetM(DZ) == m && mxGetN(DZ) == n, "Lsymb.dz size mismatch.");
  MEX_REQUIRE(mxIsSparse(DZ), "Lsymb.dz must be sparse.");
  const mxArray *PM = need_field(LS, "perm", "Missing field Lsymb.perm.");
  MEX_REQUIRE(numel(PM) == n, "Size mismatch Lsymb.perm.");
  const mxArray *FI = need_field(LS, "first", "Missing field Lsymb.first.");
  MEX_REQUIRE(numel(FI) == n, "Size mismatch Lsymb.first.");
  std::vector<sb_idx> colperm, first;
  idx_from_double(PM, colperm, 1, "Lsymb.perm");
  idx_from_double(FI, first, 1, "Lsymb.first");
  const mwIndex *dzjc = mxGetJc(DZ);
  mwSize pnnz = 0;
  for (mwSize i = 1; i <= n; i++) pnnz += dzjc[i];
  double *p = (double *)mxCalloc(pnnz ? pnnz : 1, sizeof(double));
  double *beta = (double *)mxCalloc(pnnz ? pnnz : 1, sizeof(double));
  std::vector<sb_idx> betajc(n + 1), pivperm(pnnz ? pnnz : 1);
  std::vector<double> dopiv(n ? n : 1);
  mxArray *dout = mxCreateDoubleMatrix(mxGetM(D), mxGetN(D), mxREAL);
  sb_idx nbeta = 0, npiv = 0;
  int rc = sb200_dpr1fact((sb_idx)m, (sb_idx)n, as_idx(mxGetJc(X)), as_idx(mxGetIr(X)), mxGetPr(X), mxGetPr(D), as_idx(dzjc),
                          as_idx(mxGetIr(DZ)), colperm.data(), first.data(), mxGetPr(SM), maxu, p, beta, betajc.data(),
                          pivperm.data(), dopiv.data(), mxGetPr(dout), &nbeta, &npiv);
  if (rc) { mxFree(p); mxFree(beta); mxDestroyArray(dout); sb_check(rc, "dpr1fact"); }
  const char *names[] = {"betajc", "beta", "p", "pivperm", "dopiv"};
  mxArray *Lden = mxCreateStructMatrix(1, 1, 5, names);
  // p and beta: buffers from the MEX allocator transplanted into the outputs (dpr1fact.c:762-769,806-816)
  mxArray *f = mxCreateDoubleMatrix(pnnz, 1, mxREAL);
  if (pnnz > 0) { mxFree(mxGetPr(f)); mxSetPr(f, p); } else mxFree(p);
  mxSetField(Lden, 0, "p", f);
  f = mxCreateDoubleMatrix((mwSize)npiv, 1, mxREAL);
  for (sb_idx i = 0; i < npiv; i++) mxGetPr(f)[i] = (double)pivperm[i];          // C-style, 0-based
  mxSetField(Lden, 0, "pivperm", f);
  f = mxCreateDoubleMatrix(n + 1, 1, mxREAL);
  for (mwSize i = 0; i <= n; i++) mxGetPr(f)[i] = (double)betajc[i] + 1.0;         // 1-based
  mxSetField(Lden, 0, "betajc", f);
  f = mxCreateDoubleMatrix((mwSize)nbeta, 1, mxREAL);
  if (nbeta > 0) {
    mxFree(mxGetPr(f));
    double *b2 = (double *)mxRealloc(beta, (size_t)nbeta * sizeof(double));
    if (!b2) mexErrMsgTxt("Memory allocation error");
    mxSetPr(f, b2);
  } else mxFree(beta);
  mxSetField(Lden, 0, "beta", f);
  f = mxCreateDoubleMatrix(n, 1, mxREAL);
  for (mwSize i = 0; i < n; i++) mxGetPr(f)[i] = dopiv[i];
  mxSetField(Lden, 0, "dopiv", f);
  plhs[0] = Lden;
  if (nlhs >= 2) plhs[1] = dout; else mxDestroyArray(dout);
}
