/* oracle/shim_sexp/sexp_min.c — TEST INFRASTRUCTURE ONLY: the few R API functions behind
 * oracle/shim_sexp/Rinternals.h. */
#include "Rinternals.h"

#include <stdarg.h>
#include <stdint.h>

static struct SEXPREC nil_rec = { NILSXP, 0, NULL, NULL, NULL, NULL, 0, 0 };
static struct SEXPREC na_string_rec = { CHARSXP, 2, (void*) "NA", NULL, NULL, NULL, 0, 0 };
static struct SEXPREC names_symbol_rec = { SYMSXP, 5, (void*) "names", NULL, NULL, NULL, 0, 0 };
SEXP R_NilValue = &nil_rec, R_NaString = &na_string_rec, R_NamesSymbol = &names_symbol_rec;
double R_NaReal;
char shim_last_error[1024];
jmp_buf* shim_error_jump;

static void __attribute__((constructor)) init_na(void) {
    union { double d; uint64_t u; } x;
    x.u = 0x7FF00000000007A2ull;          /* R's NA_real_: a NaN with payload 1954 */
    R_NaReal = x.d;
}
int R_IsNA(double v) {
    union { double d; uint64_t u; } x;
    x.d = v;
    return isnan(v) && (uint32_t) x.u == 1954u;
}

static SEXP fresh(SEXPTYPE type, R_xlen_t n, size_t elt) {
    SEXP s = calloc(1, sizeof *s);
    s->type = type; s->len = n;
    s->data = calloc(n > 0 ? (size_t) n : 1, elt);
    s->tag = s->prot = s->names = R_NilValue;
    return s;
}

SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n) {
    switch (type) {
    case INTSXP: case LGLSXP: return fresh(type, n, sizeof(int));
    case REALSXP: return fresh(type, n, sizeof(double));
    case STRSXP: case VECSXP: {
        SEXP s = fresh(type, n, sizeof(SEXP));
        for (R_xlen_t i = 0; i < n; ++i) ((SEXP*) s->data)[i] = type == STRSXP ? Rf_mkChar("") : R_NilValue;
        return s;
    }
    default: Rf_error("shim: allocVector of type %u", type);
    }
}
SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol) {
    SEXP s = Rf_allocVector(type, (R_xlen_t) nrow * ncol);
    s->nrow = nrow; s->ncol = ncol;
    return s;
}
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); INTEGER(s)[0] = v; return s; }
SEXP Rf_ScalarLogical(int v) { SEXP s = Rf_allocVector(LGLSXP, 1); LOGICAL(s)[0] = v; return s; }
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }
SEXP Rf_mkChar(const char* str) {
    SEXP s = calloc(1, sizeof *s);
    s->type = CHARSXP; s->len = (R_xlen_t) strlen(str); s->data = strdup(str);
    s->tag = s->prot = s->names = R_NilValue;
    return s;
}
SEXP Rf_mkString(const char* str) { SEXP s = Rf_allocVector(STRSXP, 1); SET_STRING_ELT(s, 0, Rf_mkChar(str)); return s; }
SEXP Rf_install(const char* name) { SEXP s = Rf_mkChar(name); s->type = SYMSXP; return s; }

int* INTEGER(SEXP x) { return (int*) x->data; }
int* LOGICAL(SEXP x) { return (int*) x->data; }
double* REAL(SEXP x) { return (double*) x->data; }
const char* R_CHAR(SEXP x) { return (const char*) x->data; }
SEXP STRING_ELT(SEXP x, R_xlen_t i) { return ((SEXP*) x->data)[i]; }
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP*) x->data)[i] = v; }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { return ((SEXP*) x->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP*) x->data)[i] = v; return v; }
R_xlen_t Rf_xlength(SEXP x) { return x->len; }
int Rf_length(SEXP x) { return (int) x->len; }
int TYPEOF(SEXP x) { return (int) x->type; }
int Rf_isNull(SEXP x) { return x == R_NilValue || x->type == NILSXP; }
int Rf_isLogical(SEXP x) { return x->type == LGLSXP; }
SEXP Rf_getAttrib(SEXP x, SEXP what) { return what == R_NamesSymbol ? x->names : R_NilValue; }

int Rf_asInteger(SEXP x) {
    if (x->len < 1) return NA_INTEGER;
    switch (x->type) {
    case INTSXP: case LGLSXP: return INTEGER(x)[0];
    case REALSXP: return (isnan(REAL(x)[0]) || fabs(REAL(x)[0]) > INT_MAX) ? NA_INTEGER : (int) REAL(x)[0];
    default: return NA_INTEGER;
    }
}
int Rf_asLogical(SEXP x) {
    if (x->len < 1) return NA_LOGICAL;
    switch (x->type) {
    case LGLSXP: return LOGICAL(x)[0];
    case INTSXP: return INTEGER(x)[0] == NA_INTEGER ? NA_LOGICAL : INTEGER(x)[0] != 0;
    case REALSXP: return isnan(REAL(x)[0]) ? NA_LOGICAL : REAL(x)[0] != 0;
    default: return NA_LOGICAL;
    }
}
double Rf_asReal(SEXP x) {
    if (x->len < 1) return NA_REAL;
    switch (x->type) {
    case REALSXP: return REAL(x)[0];
    case INTSXP: case LGLSXP: return INTEGER(x)[0] == NA_INTEGER ? NA_REAL : (double) INTEGER(x)[0];
    default: return NA_REAL;
    }
}
SEXP Rf_asChar(SEXP x) {
    if (x->type == CHARSXP) return x;
    if (x->type == SYMSXP) { SEXP s = Rf_mkChar((const char*) x->data); return s; }
    if (x->type == STRSXP && x->len >= 1) return STRING_ELT(x, 0);
    return R_NaString;
}

SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) {
    SEXP s = calloc(1, sizeof *s);
    s->type = EXTPTRSXP; s->data = p; s->tag = tag; s->prot = prot; s->names = R_NilValue;
    return s;
}
void* R_ExternalPtrAddr(SEXP x) { return x->data; }
SEXP R_ExternalPtrTag(SEXP x) { return x->tag; }
void R_ClearExternalPtr(SEXP x) { x->data = NULL; }
void R_RegisterCFinalizerEx(SEXP x, R_CFinalizer_t fun, int onexit) { (void) x; (void) fun; (void) onexit; }
char* R_alloc(size_t n, int size) { return (char*) calloc(n ? n : 1, (size_t) size); }

void Rprintf(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
}
void Rf_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(shim_last_error, sizeof shim_last_error, fmt, ap);
    va_end(ap);
    if (shim_error_jump) longjmp(*shim_error_jump, 1);
    fprintf(stderr, "Rf_error: %s\n", shim_last_error);
    abort();
}

/* ---- helpers for the tests (ctypes): build arguments, guard a call against Rf_error */
SEXP shim_int_vector(R_xlen_t n, const int* v) { SEXP s = Rf_allocVector(INTSXP, n); if (n) memcpy(s->data, v, sizeof(int) * (size_t) n); return s; }
SEXP shim_real_vector(R_xlen_t n, const double* v) { SEXP s = Rf_allocVector(REALSXP, n); if (n) memcpy(s->data, v, sizeof(double) * (size_t) n); return s; }
SEXP shim_logical(int v) { return Rf_ScalarLogical(v); }
SEXP shim_string(const char* s) { return s ? Rf_mkString(s) : R_NilValue; }
SEXP shim_nil(void) { return R_NilValue; }
SEXP shim_extptr(void* p) { return R_MakeExternalPtr(p, R_NilValue, R_NilValue); }
R_xlen_t shim_len(SEXP s) { return s->len; }
int shim_type(SEXP s) { return (int) s->type; }
int shim_nrow(SEXP s) { return (int) s->nrow; }
int shim_ncol(SEXP s) { return (int) s->ncol; }
void* shim_data(SEXP s) { return s->data; }
const char* shim_error_text(void) { return shim_last_error; }
