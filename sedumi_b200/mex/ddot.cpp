// ddotx = ddot(d,X,blkstart[,Xblkjc])   d[k]'*x[k] per Lorentz block (ddot.c:48-56 signature, :165-308)
#include "mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  MEX_REQUIRE(nrhs >= 3, "ddot requires more input arguments.");
  MEX_REQUIRE(nlhs <= 1, "ddot generates less output arguments.");
  const mxArray *D = prhs[0], *X = prhs[1], *BLK = prhs[2];
  MEX_REQUIRE(numel(BLK) >= 1, "blkstart size mismatch.");
  std::vector<sb_idx> bs;
  idx_from_double(BLK, bs, 1, "blkstart");
  sb_idx nblk = (sb_idx)bs.size() - 1;
  const double *d = mxGetPr(D);
  sb_idx qDim = (sb_idx)numel(D);
  if (qDim != bs[nblk] - bs[0]) {
    MEX_REQUIRE(qDim >= bs[nblk], "d size mismatch.");
    d += bs[0];                                          // point to the Lorentz norm-bound part
    qDim = bs[nblk] - bs[0];
  }
  mwSize nrows = mxGetM(X), m = mxGetN(X);
  if (!mxIsSparse(X)) {
    const double *x = mxGetPr(X);
    if ((sb_idx)nrows != qDim) {
      if ((sb_idx)nrows < bs[nblk]) { MEX_REQUIRE((sb_idx)nrows == nblk + qDim, "X size mismatch"); x += nblk; }
      else x += bs[0];
    }
    plhs[0] = mxCreateDoubleMatrix((mwSize)nblk, m, mxREAL);
    std::vector<sb_idx> rel(nblk + 1);
    for (sb_idx k = 0; k <= nblk; k++) rel[k] = bs[k] - bs[0];
    int rc = sb200_ddot_dense(nblk, rel.data(), d, x, (sb_idx)nrows, (sb_idx)m, mxGetPr(plhs[0]));
    if (rc) { mxDestroyArray(plhs[0]); plhs[0] = NULL; sb_check(rc, "ddot"); }
    return;
  }
  MEX_REQUIRE((sb_idx)nrows >= bs[nblk], "X size mismatch");
  MEX_REQUIRE(nrhs >= 4, "ddot with sparse X requires more input arguments.");
  const mxArray *XJC = prhs[3];
  MEX_REQUIRE(mxGetM(XJC) == m && mxGetN(XJC) >= 3, "Xjc size mismatch");
  const double *xjcPr = mxGetPr(XJC) + m;                // Xjc(:,2) and Xjc(:,3)
  std::vector<sb_idx> xlo(m), xhi(m);
  mwSize maxnnz = 0;
  for (mwSize j = 0; j < m; j++) {
    xlo[j] = (sb_idx)xjcPr[j]; xhi[j] = (sb_idx)xjcPr[m + j];
    MEX_REQUIRE(xhi[j] >= xlo[j], "Xjc must be increasing");
    maxnnz += (mwSize)(xhi[j] - xlo[j]);
  }
  mxArray *out = mxCreateSparse((mwSize)nblk, m, maxnnz > 0 ? maxnnz : 1, mxREAL);
  sb_idx nnz = 0;
  int rc = sb200_ddot_sparse(nblk, bs.data(), d, (sb_idx)m, xlo.data(), xhi.data(), as_idx(mxGetIr(X)), mxGetPr(X),
                             (sb_idx *)mxGetJc(out), (sb_idx *)mxGetIr(out), mxGetPr(out), &nnz);
  if (rc) { mxDestroyArray(out); sb_check(rc, "ddot"); }
  mwSize keep = nnz > 0 ? (mwSize)nnz : 1;                // shrink like the reference (ddot.c:288-297)
  mwIndex *ir = (mwIndex *)mxRealloc(mxGetIr(out), keep * sizeof(mwIndex));
  double *pr = (double *)mxRealloc(mxGetPr(out), keep * sizeof(double));
  if (!ir || !pr) mexErrMsgTxt("Memory allocation error");
  mxSetIr(out, ir); mxSetPr(out, pr); mxSetNzmax(out, keep);
  plhs[0] = out;
}
