// y = fwblkslv(L,b[,ysymb])   y = L.L \ b(L.perm,:)   (fwblkslv.c:46-53 signature, :193-320)
#include "mex_common.h"

static void get_L(const mxArray *L_IN, mwSize &m, std::vector<sb_idx> &perm, std::vector<sb_idx> &xsuper,
                  const mxArray *&LL) {
  MEX_REQUIRE(mxIsStruct(L_IN), "Parameter `L' should be a structure.");
  const mxArray *f = need_field(L_IN, "perm", "Missing field L.perm.");
  m = numel(f);
  idx_from_double(f, perm, 1, "L.perm");
  LL = need_field(L_IN, "L", "Missing field L.L.");
  MEX_REQUIRE(mxGetM(LL) == m && mxGetN(LL) == m, "Size L.L mismatch.");
  MEX_REQUIRE(mxIsSparse(LL), "L.L should be sparse.");
  f = need_field(L_IN, "xsuper", "Missing field L.xsuper.");
  MEX_REQUIRE(numel(f) >= 1 && numel(f) - 1 <= m, "Size L.xsuper mismatch.");
  idx_from_double(f, xsuper, 1, "L.xsuper");
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 2, "fwblkslv requires more input arguments.");
  MEX_REQUIRE(nlhs <= 1, "fwblkslv generates only 1 output argument.");
  mwSize m;
  std::vector<sb_idx> perm, xsuper;
  const mxArray *LL;
  get_L(prhs[0], m, perm, xsuper, LL);
  const mxArray *B = prhs[1];
  MEX_REQUIRE(mxGetM(B) == m, "Size mismatch b.");
  mwSize n = mxGetN(B);
  sb_idx nsuper = (sb_idx)xsuper.size() - 1;
  if (!mxIsSparse(B)) {
    plhs[0] = mxCreateDoubleMatrix(m, n, mxREAL);
    int rc = sb200_fwblkslv((sb_idx)m, nsuper, xsuper.data(), as_idx(mxGetJc(LL)), as_idx(mxGetIr(LL)), mxGetPr(LL),
                            perm.data(), mxGetPr(B), mxGetPr(plhs[0]), (sb_idx)n);
    if (rc) { mxDestroyArray(plhs[0]); plhs[0] = NULL; sb_check(rc, "fwblkslv"); }
    return;
  }
  MEX_REQUIRE(nrhs >= 3, "fwblkslv requires more inputs in case of sparse b.");
  const mxArray *Y = prhs[2];
  MEX_REQUIRE(mxGetM(Y) == m && mxGetN(Y) == n, "Size mismatch y.");
  MEX_REQUIRE(mxIsSparse(Y), "y should be sparse.");
  const mwIndex *yjc = mxGetJc(Y), *yir = mxGetIr(Y);
  plhs[0] = mxCreateSparse(m, n, yjc[n], mxREAL);
  memcpy(mxGetJc(plhs[0]), yjc, (n + 1) * sizeof(mwIndex));
  memcpy(mxGetIr(plhs[0]), yir, yjc[n] * sizeof(mwIndex));
  int rc = sb200_fwblkslv_sparse((sb_idx)m, nsuper, xsuper.data(), as_idx(mxGetJc(LL)), as_idx(mxGetIr(LL)), mxGetPr(LL),
                                 perm.data(), (sb_idx)n, as_idx(mxGetJc(B)), as_idx(mxGetIr(B)), mxGetPr(B),
                                 as_idx(yjc), as_idx(yir), mxGetPr(plhs[0]));
  if (rc) { mxDestroyArray(plhs[0]); plhs[0] = NULL; sb_check(rc, "fwblkslv"); }
}
