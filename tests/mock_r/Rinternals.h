/* Stand-in for R's Rinternals.h: just enough declarations for a syntax and type check of
 * integration/scasa_r.c against include/scasa_cuda.h (R is not installed here).  Not R's header. */
#pragma once
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned char Rbyte;
typedef void (*R_CFinalizer_t)(SEXP);
typedef enum { FALSE = 0, TRUE } Rboolean;
#define REALSXP 14
#define INTSXP 13
extern SEXP R_NilValue;
int *INTEGER(SEXP);
double *REAL(SEXP);
Rbyte *RAW(SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
R_xlen_t XLENGTH(SEXP);
int asInteger(SEXP);
int ncols(SEXP);
SEXP allocMatrix(unsigned int, int, int);
SEXP PROTECT(SEXP);
void UNPROTECT(int);
void *R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
SEXP R_MakeExternalPtr(void *, SEXP, SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
char *R_alloc(size_t, int);
void error(const char *, ...) __attribute__((noreturn));
#define VECSXP 19
#define STRSXP 16
double asReal(SEXP);
SEXP ScalarLogical(int);
SEXP allocVector(unsigned int, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
const char *CHAR(SEXP);
