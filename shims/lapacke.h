/*
 * lapacke.h -- ENVIRONMENT SHIM (see shims/README.md).
 *
 * LAPACKE declaration shim for compiling the unmodified reference sources into
 * oracle/_ref/.  Maps each LAPACKE routine the reference calls onto the
 * `scipy_LAPACKE_*` symbol of the scipy-bundled LP64 OpenBLAS.  None of these
 * routines are on the Chebyshev-filter hot path; they are only reached by the
 * full-program oracle (subspace eigenproblem, mixing, MLFF, ...).
 */
#ifndef ORACLE_SHIM_LAPACKE_H
#define ORACLE_SHIM_LAPACKE_H

#include <complex.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
typedef int lapack_int;
typedef double _Complex lapack_complex_double;

#define LAPACKE_dsyevd scipy_LAPACKE_dsyevd
#define LAPACKE_dgelsd scipy_LAPACKE_dgelsd
#define LAPACKE_zhegvd scipy_LAPACKE_zhegvd
#define LAPACKE_dsygvd scipy_LAPACKE_dsygvd
#define LAPACKE_dsterf scipy_LAPACKE_dsterf
#define LAPACKE_dsyev  scipy_LAPACKE_dsyev
#define LAPACKE_dgesvd scipy_LAPACKE_dgesvd
#define LAPACKE_dgeev  scipy_LAPACKE_dgeev
#define LAPACKE_dpotrf scipy_LAPACKE_dpotrf
#define LAPACKE_zpotrf scipy_LAPACKE_zpotrf
#define LAPACKE_zheev  scipy_LAPACKE_zheev
#define LAPACKE_zggev  scipy_LAPACKE_zggev
#define LAPACKE_dsysv  scipy_LAPACKE_dsysv
#define LAPACKE_dlange scipy_LAPACKE_dlange
#define LAPACKE_dggev  scipy_LAPACKE_dggev
#define LAPACKE_dgetrf scipy_LAPACKE_dgetrf
#define LAPACKE_dgesv  scipy_LAPACKE_dgesv
#define LAPACKE_dgecon scipy_LAPACKE_dgecon

lapack_int LAPACKE_dsyevd(int layout, char jobz, char uplo, lapack_int n, double *a,
                          lapack_int lda, double *w);
lapack_int LAPACKE_dgelsd(int layout, lapack_int m, lapack_int n, lapack_int nrhs, double *a,
                          lapack_int lda, double *b, lapack_int ldb, double *s, double rcond,
                          lapack_int *rank);
lapack_int LAPACKE_zhegvd(int layout, lapack_int itype, char jobz, char uplo, lapack_int n,
                          lapack_complex_double *a, lapack_int lda, lapack_complex_double *b,
                          lapack_int ldb, double *w);
lapack_int LAPACKE_dsygvd(int layout, lapack_int itype, char jobz, char uplo, lapack_int n,
                          double *a, lapack_int lda, double *b, lapack_int ldb, double *w);
lapack_int LAPACKE_dsterf(lapack_int n, double *d, double *e);
lapack_int LAPACKE_dsyev(int layout, char jobz, char uplo, lapack_int n, double *a,
                         lapack_int lda, double *w);
lapack_int LAPACKE_dgesvd(int layout, char jobu, char jobvt, lapack_int m, lapack_int n,
                          double *a, lapack_int lda, double *s, double *u, lapack_int ldu,
                          double *vt, lapack_int ldvt, double *superb);
lapack_int LAPACKE_dgeev(int layout, char jobvl, char jobvr, lapack_int n, double *a,
                         lapack_int lda, double *wr, double *wi, double *vl, lapack_int ldvl,
                         double *vr, lapack_int ldvr);
lapack_int LAPACKE_dpotrf(int layout, char uplo, lapack_int n, double *a, lapack_int lda);
lapack_int LAPACKE_zpotrf(int layout, char uplo, lapack_int n, lapack_complex_double *a,
                          lapack_int lda);
lapack_int LAPACKE_zheev(int layout, char jobz, char uplo, lapack_int n,
                         lapack_complex_double *a, lapack_int lda, double *w);
lapack_int LAPACKE_zggev(int layout, char jobvl, char jobvr, lapack_int n,
                         lapack_complex_double *a, lapack_int lda, lapack_complex_double *b,
                         lapack_int ldb, lapack_complex_double *alpha,
                         lapack_complex_double *beta, lapack_complex_double *vl,
                         lapack_int ldvl, lapack_complex_double *vr, lapack_int ldvr);
lapack_int LAPACKE_dsysv(int layout, char uplo, lapack_int n, lapack_int nrhs, double *a,
                         lapack_int lda, lapack_int *ipiv, double *b, lapack_int ldb);
double LAPACKE_dlange(int layout, char norm, lapack_int m, lapack_int n, const double *a,
                      lapack_int lda);
lapack_int LAPACKE_dggev(int layout, char jobvl, char jobvr, lapack_int n, double *a,
                         lapack_int lda, double *b, lapack_int ldb, double *alphar,
                         double *alphai, double *beta, double *vl, lapack_int ldvl, double *vr,
                         lapack_int ldvr);
lapack_int LAPACKE_dgetrf(int layout, lapack_int m, lapack_int n, double *a, lapack_int lda,
                          lapack_int *ipiv);
lapack_int LAPACKE_dgesv(int layout, lapack_int n, lapack_int nrhs, double *a, lapack_int lda,
                         lapack_int *ipiv, double *b, lapack_int ldb);
lapack_int LAPACKE_dgecon(int layout, char norm, lapack_int n, const double *a, lapack_int lda,
                          double anorm, double *rcond);

#ifdef __cplusplus
}
#endif
#endif
