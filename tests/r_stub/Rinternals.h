/* Declarations-only stand-in for the part of R's C API that resnmtf_b200/r/r_shim.c uses, so that the shim is at
 * least COMPILED in this image (no R here): tests/test_abi.py builds r_shim.c against it with -Wall -Werror.
 * Signatures follow R's Rinternals.h (R >= 4.0); nothing here is linked or run. */
#ifndef RESNMTF_R_STUB_RINTERNALS_H
#define RESNMTF_R_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;
typedef unsigned int SEXPTYPE;
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
extern SEXP R_NilValue;
extern SEXP R_NamesSymbol;
extern int R_NaInt;
#define NA_INTEGER R_NaInt
#ifdef __GNUC__
#define RSTUB_NORETURN __attribute__((noreturn))
#else
#define RSTUB_NORETURN
#endif
void RSTUB_NORETURN Rf_error(const char*, ...);
void Rf_warning(const char*, ...);
int LENGTH(SEXP);
R_xlen_t XLENGTH(SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
const char* CHAR(SEXP);
double* REAL(SEXP);
int* INTEGER(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
Rboolean Rf_isReal(SEXP);
Rboolean Rf_isMatrix(SEXP);
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_mkNamed(SEXPTYPE, const char**);
SEXP Rf_mkString(const char*);
SEXP Rf_ScalarReal(double);
SEXP Rf_ScalarInteger(int);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
char* R_alloc(size_t, int);
typedef void (*R_CFinalizer_t)(SEXP);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
void R_CheckUserInterrupt(void);
#endif
