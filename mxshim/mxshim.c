/* mxshim.c -- implementation of the MEX API subset declared in mex.h.
 * Test/bench infrastructure and the host side of the drop-in boundary when no
 * MATLAB/Octave is present.  See mex.h for scope. */
#include "mex.h"
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

struct mxArray_tag {
  int      classid;          /* mxShimClass */
  mwSize   m, n;
  double  *pr;
  double  *pi;               /* never populated on this path; kept for API shape */
  mwIndex *ir, *jc;
  mwSize   nzmax;
  int      nfields;
  char   **fieldnames;
  mxArray **fields;
};

static __thread jmp_buf *g_jmp = NULL;
static __thread char g_err[1024];

/* ------------------------------------------------------------ allocator */
void *mxCalloc(size_t n, size_t size) {
  if (n == 0 || size == 0) { n = 1; size = 1; }
  return calloc(n, size);
}
void *mxMalloc(size_t n) { return malloc(n ? n : 1); }
void *mxRealloc(void *p, size_t size) { return realloc(p, size ? size : 1); }
void  mxFree(void *p) { free(p); }

/* ------------------------------------------------------------ creation */
static mxArray *new_array(int classid, mwSize m, mwSize n) {
  mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
  a->classid = classid; a->m = m; a->n = n;
  return a;
}

mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity flag) {
  mxArray *a = new_array(mxSHIM_DOUBLE, m, n);
  a->pr = (double *)mxCalloc(m * n, sizeof(double));
  if (flag == mxCOMPLEX) a->pi = (double *)mxCalloc(m * n, sizeof(double));
  return a;
}

mxArray *mxCreateDoubleScalar(double v) {
  mxArray *a = mxCreateDoubleMatrix(1, 1, mxREAL);
  a->pr[0] = v;
  return a;
}

mxArray *mxCreateSparse(mwSize m, mwSize n, mwSize nzmax, mxComplexity flag) {
  mxArray *a = new_array(mxSHIM_SPARSE, m, n);
  if (nzmax < 1) nzmax = 1;                 /* MATLAB also forces nzmax >= 1 */
  a->nzmax = nzmax;
  a->pr = (double *)mxCalloc(nzmax, sizeof(double));
  a->ir = (mwIndex *)mxCalloc(nzmax, sizeof(mwIndex));
  a->jc = (mwIndex *)mxCalloc(n + 1, sizeof(mwIndex));
  if (flag == mxCOMPLEX) a->pi = (double *)mxCalloc(nzmax, sizeof(double));
  return a;
}

mxArray *mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char **fieldnames) {
  mxArray *a = new_array(mxSHIM_STRUCT, m, n);
  int i;
  a->nfields = nfields;
  a->fieldnames = (char **)calloc(nfields > 0 ? nfields : 1, sizeof(char *));
  a->fields = (mxArray **)calloc(nfields > 0 ? nfields : 1, sizeof(mxArray *));
  for (i = 0; i < nfields; i++) a->fieldnames[i] = strdup(fieldnames[i]);
  return a;
}

mxArray *mxDuplicateArray(const mxArray *s) {
  mxArray *a;
  int i;
  if (!s) return NULL;
  a = new_array(s->classid, s->m, s->n);
  switch (s->classid) {
  case mxSHIM_DOUBLE:
    a->pr = (double *)mxCalloc(s->m * s->n, sizeof(double));
    if (s->pr) memcpy(a->pr, s->pr, s->m * s->n * sizeof(double));
    break;
  case mxSHIM_SPARSE:
    a->nzmax = s->nzmax;
    a->pr = (double *)mxCalloc(s->nzmax, sizeof(double));
    a->ir = (mwIndex *)mxCalloc(s->nzmax, sizeof(mwIndex));
    a->jc = (mwIndex *)mxCalloc(s->n + 1, sizeof(mwIndex));
    memcpy(a->jc, s->jc, (s->n + 1) * sizeof(mwIndex));
    {
      mwSize nnz = s->jc[s->n];
      if (nnz > s->nzmax) nnz = s->nzmax;
      if (s->pr) memcpy(a->pr, s->pr, nnz * sizeof(double));
      memcpy(a->ir, s->ir, nnz * sizeof(mwIndex));
    }
    break;
  case mxSHIM_STRUCT:
    a->nfields = s->nfields;
    a->fieldnames = (char **)calloc(s->nfields > 0 ? s->nfields : 1, sizeof(char *));
    a->fields = (mxArray **)calloc(s->nfields > 0 ? s->nfields : 1, sizeof(mxArray *));
    for (i = 0; i < s->nfields; i++) {
      a->fieldnames[i] = strdup(s->fieldnames[i]);
      a->fields[i] = mxDuplicateArray(s->fields[i]);
    }
    break;
  }
  return a;
}

void mxDestroyArray(mxArray *a) {
  int i;
  if (!a) return;
  free(a->pr); free(a->pi); free(a->ir); free(a->jc);
  if (a->classid == mxSHIM_STRUCT) {
    for (i = 0; i < a->nfields; i++) {
      free(a->fieldnames[i]);
      mxDestroyArray(a->fields[i]);
    }
    free(a->fieldnames); free(a->fields);
  }
  free(a);
}

/* ------------------------------------------------------------ queries */
double  *mxGetPr(const mxArray *a) { return a->pr; }
double  *mxGetPi(const mxArray *a) { return a->pi; }
mwIndex *mxGetJc(const mxArray *a) { return a->jc; }
mwIndex *mxGetIr(const mxArray *a) { return a->ir; }
mwSize   mxGetM(const mxArray *a) { return a->m; }
mwSize   mxGetN(const mxArray *a) { return a->n; }
mwSize   mxGetNzmax(const mxArray *a) { return a->nzmax; }
mwSize   mxGetNumberOfElements(const mxArray *a) { return a->m * a->n; }
bool     mxIsSparse(const mxArray *a) { return a->classid == mxSHIM_SPARSE; }
bool     mxIsStruct(const mxArray *a) { return a->classid == mxSHIM_STRUCT; }
bool     mxIsDouble(const mxArray *a) { return a->classid != mxSHIM_STRUCT; }
bool     mxIsEmpty(const mxArray *a) { return a->m == 0 || a->n == 0; }
bool     mxIsComplex(const mxArray *a) { return a->pi != NULL; }
int      mxshim_class(const mxArray *a) { return a->classid; }

double mxGetScalar(const mxArray *a) {
  if (!a || !a->pr) return 0.0;
  if (a->classid == mxSHIM_SPARSE && a->jc[a->n] == 0) return 0.0;
  if (a->m * a->n == 0) return 0.0;          /* MATLAB: undefined; be benign */
  return a->pr[0];
}

/* ------------------------------------------------------------ struct fields */
static int field_index(const mxArray *a, const char *name) {
  int i;
  if (!a || a->classid != mxSHIM_STRUCT) return -1;
  for (i = 0; i < a->nfields; i++)
    if (strcmp(a->fieldnames[i], name) == 0) return i;
  return -1;
}
mxArray *mxGetField(const mxArray *a, mwIndex index, const char *fieldname) {
  int i = field_index(a, fieldname);
  (void)index;
  return i < 0 ? NULL : a->fields[i];
}
int mxAddField(mxArray *a, const char *fieldname) {
  int i = field_index(a, fieldname);
  if (i >= 0) return i;
  a->fieldnames = (char **)realloc(a->fieldnames, (a->nfields + 1) * sizeof(char *));
  a->fields = (mxArray **)realloc(a->fields, (a->nfields + 1) * sizeof(mxArray *));
  a->fieldnames[a->nfields] = strdup(fieldname);
  a->fields[a->nfields] = NULL;
  return a->nfields++;
}
void mxSetField(mxArray *a, mwIndex index, const char *fieldname, mxArray *value) {
  int i = field_index(a, fieldname);
  (void)index;
  if (i < 0) i = mxAddField(a, fieldname);
  a->fields[i] = value;       /* like MATLAB: previous content is NOT freed */
}
int mxGetNumberOfFields(const mxArray *a) { return a->classid == mxSHIM_STRUCT ? a->nfields : 0; }
const char *mxGetFieldNameByNumber(const mxArray *a, int n) {
  return (a->classid == mxSHIM_STRUCT && n >= 0 && n < a->nfields) ? a->fieldnames[n] : NULL;
}

/* ------------------------------------------------------------ setters */
void mxSetPr(mxArray *a, double *pr) { a->pr = pr; }
void mxSetIr(mxArray *a, mwIndex *ir) { a->ir = ir; }
void mxSetJc(mxArray *a, mwIndex *jc) { a->jc = jc; }
void mxSetM(mxArray *a, mwSize m) { a->m = m; }
void mxSetN(mxArray *a, mwSize n) { a->n = n; }
void mxSetNzmax(mxArray *a, mwSize nzmax) { a->nzmax = nzmax; }

/* ------------------------------------------------------------ errors */
static void raise_error(void) {
  if (g_jmp) longjmp(*g_jmp, 1);
  fprintf(stderr, "mxshim: error outside mxshim_call: %s\n", g_err);
  abort();
}
void mexErrMsgTxt(const char *msg) {
  snprintf(g_err, sizeof g_err, "%s", msg ? msg : "(null)");
  raise_error();
}
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...) {
  char buf[900];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  snprintf(g_err, sizeof g_err, "%s: %s", id ? id : "", buf);
  raise_error();
}
void mexWarnMsgTxt(const char *msg) { fprintf(stderr, "mex warning: %s\n", msg); }
int mexPrintf(const char *fmt, ...) {
  int r;
  va_list ap;
  va_start(ap, fmt);
  r = vprintf(fmt, ap);
  va_end(ap);
  return r;
}
void mxshim_assert_fail(const char *expr, const char *msg, const char *file, int line) {
  snprintf(g_err, sizeof g_err, "mxAssert(%s) failed at %s:%d: %s", expr, file, line, msg ? msg : "");
  raise_error();
}

int mxshim_call(mxshim_mexfn fn, int nlhs, mxArray **plhs, int nrhs, const mxArray **prhs) {
  jmp_buf env;
  jmp_buf *prev = g_jmp;
  int rc = 0;
  g_err[0] = 0;
  g_jmp = &env;
  if (setjmp(env) == 0) fn(nlhs, plhs, nrhs, prhs);
  else rc = 1;
  g_jmp = prev;
  return rc;
}
const char *mxshim_last_error(void) { return g_err; }
