/* blas1.c -- TEST INFRASTRUCTURE (oracle build only).
 * The five reference-BLAS level-1 routines SeDuMi's C calls (Fortran calling
 * convention, standard semantics: idamax_ returns a 1-BASED index of the first
 * element of maximum |x|).  Plain sequential loops, i.e. netlib reference BLAS
 * summation order, so the oracle is reproducible bit-for-bit run to run.
 * Call sites in the reference: sdmauxRdot.c:47,57; sdmauxScalarmul.c:44-66;
 * blkchol2.c:56,69. */
#include <math.h>
typedef int blasint;

double ddot_(blasint *n, double *x, blasint *incx, double *y, blasint *incy) {
  double s = 0.0;
  blasint i, ix = *incx, iy = *incy;
  for (i = 0; i < *n; i++) s += x[i * ix] * y[i * iy];
  return s;
}
int daxpy_(blasint *n, double *alpha, double *x, blasint *incx, double *y, blasint *incy) {
  blasint i, ix = *incx, iy = *incy;
  double a = *alpha;
  if (a == 0.0) return 0;
  for (i = 0; i < *n; i++) y[i * iy] += a * x[i * ix];
  return 0;
}
int dscal_(blasint *n, double *alpha, double *x, blasint *incx) {
  blasint i, ix = *incx;
  double a = *alpha;
  for (i = 0; i < *n; i++) x[i * ix] *= a;
  return 0;
}
int dcopy_(blasint *n, double *x, blasint *incx, double *y, blasint *incy) {
  blasint i, ix = *incx, iy = *incy;
  for (i = 0; i < *n; i++) y[i * iy] = x[i * ix];
  return 0;
}
blasint idamax_(blasint *n, double *x, blasint *incx) {
  blasint i, imax = 0, ix = *incx;
  double vmax;
  if (*n < 1) return 0;
  vmax = fabs(x[0]);
  for (i = 1; i < *n; i++) {
    double v = fabs(x[i * ix]);
    if (v > vmax) { vmax = v; imax = i; }
  }
  return imax + 1;      /* Fortran: 1-based */
}
