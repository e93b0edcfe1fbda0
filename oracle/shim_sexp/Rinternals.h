/* oracle/shim_sexp/Rinternals.h — TEST INFRASTRUCTURE ONLY.
 *
 * A minimal stand-in for R's C API (R.h / Rinternals.h / Rdefines): just enough SEXP machinery
 * — integer / real / logical / string vectors, matrices, external pointers, symbols, NA values —
 * for the hot-path `SXP_*` entry points of the reference's UNMODIFIED src/r-wrappers.c to compile
 * and run without R (R is not installed in this image).  Everything is heap-allocated and never
 * collected; PROTECT / UNPROTECT are no-ops.  Used by oracle/Makefile `wrappers` and
 * tests/test_gpu_wrappers.py only. */
#ifndef GSB_SHIM_RINTERNALS_H
#define GSB_SHIM_RINTERNALS_H

#include <limits.h>
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef ptrdiff_t R_xlen_t;
typedef struct SEXPREC* SEXP;
typedef unsigned int SEXPTYPE;
enum { NILSXP = 0, SYMSXP = 1, CHARSXP = 9, LGLSXP = 10, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, EXTPTRSXP = 22 };

struct SEXPREC {
    SEXPTYPE type;
    R_xlen_t len;
    void* data;          /* int* / double* / SEXP* / char* / the external pointer */
    SEXP tag, prot;      /* external pointers */
    SEXP names;
    R_xlen_t nrow, ncol; /* matrices */
};

#define NA_INTEGER INT_MIN
#define NA_LOGICAL INT_MIN
extern double R_NaReal;
#define NA_REAL R_NaReal
int R_IsNA(double x);
#define ISNA(x) R_IsNA(x)
extern SEXP R_NilValue, R_NaString, R_NamesSymbol;
#define NA_STRING R_NaString
#define TRUE 1
#define FALSE 0

SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n);
SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol);
SEXP Rf_ScalarInteger(int v);
SEXP Rf_ScalarReal(double v);
SEXP Rf_ScalarLogical(int v);
SEXP Rf_mkChar(const char* s);
SEXP Rf_mkString(const char* s);
SEXP Rf_install(const char* name);
int* INTEGER(SEXP x);
int* LOGICAL(SEXP x);
double* REAL(SEXP x);
const char* R_CHAR(SEXP x);
#define CHAR(x) R_CHAR(x)
SEXP STRING_ELT(SEXP x, R_xlen_t i);
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
R_xlen_t Rf_xlength(SEXP x);
int Rf_length(SEXP x);
int TYPEOF(SEXP x);
int Rf_asInteger(SEXP x);
int Rf_asLogical(SEXP x);
double Rf_asReal(SEXP x);
SEXP Rf_asChar(SEXP x);
int Rf_isNull(SEXP x);
int Rf_isLogical(SEXP x);
SEXP Rf_getAttrib(SEXP x, SEXP what);
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot);
void* R_ExternalPtrAddr(SEXP x);
SEXP R_ExternalPtrTag(SEXP x);
void R_ClearExternalPtr(SEXP x);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP x, R_CFinalizer_t fun, int onexit);
char* R_alloc(size_t n, int size);
void Rprintf(const char* fmt, ...);
void Rf_error(const char* fmt, ...) __attribute__((noreturn));
/* what the last Rf_error said; tests install a jump target to survive it */
extern char shim_last_error[1024];
#include <setjmp.h>
extern jmp_buf* shim_error_jump;

#define PROTECT(x) (x)
#define UNPROTECT(n) ((void) (n))
#define R_Calloc(n, t) ((t*) calloc((size_t) (n), sizeof(t)))
#define R_Free(p) (free(p), (p) = NULL)

/* the unprefixed names r-wrappers.c uses (R without R_NO_REMAP) */
#define allocVector Rf_allocVector
#define allocMatrix Rf_allocMatrix
#define ScalarInteger Rf_ScalarInteger
#define ScalarReal Rf_ScalarReal
#define ScalarLogical Rf_ScalarLogical
#define mkChar Rf_mkChar
#define mkString Rf_mkString
#define xlength Rf_xlength
#define length Rf_length
#define asInteger Rf_asInteger
#define asLogical Rf_asLogical
#define asReal Rf_asReal
#define asChar Rf_asChar
#define isNull Rf_isNull
#define isLogical Rf_isLogical
#define getAttrib Rf_getAttrib
#define error Rf_error

#endif
