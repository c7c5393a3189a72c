/* f77blas.h -- declarations of the five BLAS level-1 routines SeDuMi's C code
 * calls under -DOCTAVE (blksdp.h:40-42: FORT(x) -> BLASFUNC(x); call sites
 * sdmauxRdot.c:47,57, sdmauxScalarmul.c:44-66, blkchol2.c:56,69).
 * Provider for the oracle build: oracle/blas1.c. */
#ifndef MXSHIM_F77BLAS_H
#define MXSHIM_F77BLAS_H
#ifdef __cplusplus
extern "C" {
#endif
typedef int blasint;
#define BLASFUNC(x) x##_
double ddot_(blasint *n, double *x, blasint *incx, double *y, blasint *incy);
int    daxpy_(blasint *n, double *alpha, double *x, blasint *incx, double *y, blasint *incy);
int    dscal_(blasint *n, double *alpha, double *x, blasint *incx);
int    dcopy_(blasint *n, double *x, blasint *incx, double *y, blasint *incy);
blasint idamax_(blasint *n, double *x, blasint *incx);
#ifdef __cplusplus
}
#endif
#endif
