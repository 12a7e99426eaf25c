/*
 * cblas.h -- ENVIRONMENT SHIM (see shims/README.md).
 *
 * Minimal CBLAS declaration shim used when the unmodified reference sources
 * under /root/reference/src are compiled into oracle/_ref/.  The image has no
 * system BLAS headers; the only BLAS available is the LP64 OpenBLAS bundled
 * with scipy, which exports every entry point with a `scipy_` prefix.  Each
 * CBLAS name the reference uses is therefore mapped onto that symbol.
 *
 * Nothing under sparc_b200/ (the product) includes this file.
 */
#ifndef ORACLE_SHIM_CBLAS_H
#define ORACLE_SHIM_CBLAS_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_LAYOUT;
typedef CBLAS_LAYOUT CBLAS_ORDER;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
typedef enum { CblasUpper = 121, CblasLower = 122 } CBLAS_UPLO;
typedef enum { CblasNonUnit = 131, CblasUnit = 132 } CBLAS_DIAG;
typedef enum { CblasLeft = 141, CblasRight = 142 } CBLAS_SIDE;

#ifndef MKL_INT
#define MKL_INT int
#endif

#define cblas_dgemm       scipy_cblas_dgemm
#define cblas_zgemm       scipy_cblas_zgemm
#define cblas_dgemv       scipy_cblas_dgemv
#define cblas_zgemv       scipy_cblas_zgemv
#define cblas_dscal       scipy_cblas_dscal
#define cblas_zscal       scipy_cblas_zscal
#define cblas_dcopy       scipy_cblas_dcopy
#define cblas_zcopy       scipy_cblas_zcopy
#define cblas_dger        scipy_cblas_dger
#define cblas_ddot        scipy_cblas_ddot
#define cblas_dtrsm       scipy_cblas_dtrsm
#define cblas_ztrsm       scipy_cblas_ztrsm
#define cblas_dsyrk       scipy_cblas_dsyrk
#define cblas_dgemm_batch scipy_cblas_dgemm_batch
#define cblas_zgemm_batch scipy_cblas_zgemm_batch

void cblas_dgemm(CBLAS_LAYOUT, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int M, int N, int K,
                 double alpha, const double *A, int lda, const double *B, int ldb,
                 double beta, double *C, int ldc);
void cblas_zgemm(CBLAS_LAYOUT, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int M, int N, int K,
                 const void *alpha, const void *A, int lda, const void *B, int ldb,
                 const void *beta, void *C, int ldc);
void cblas_dgemv(CBLAS_LAYOUT, CBLAS_TRANSPOSE, int M, int N, double alpha, const double *A,
                 int lda, const double *X, int incX, double beta, double *Y, int incY);
void cblas_zgemv(CBLAS_LAYOUT, CBLAS_TRANSPOSE, int M, int N, const void *alpha, const void *A,
                 int lda, const void *X, int incX, const void *beta, void *Y, int incY);
void cblas_dscal(int N, double alpha, double *X, int incX);
void cblas_zscal(int N, const void *alpha, void *X, int incX);
void cblas_dcopy(int N, const double *X, int incX, double *Y, int incY);
void cblas_zcopy(int N, const void *X, int incX, void *Y, int incY);
void cblas_dger(CBLAS_LAYOUT, int M, int N, double alpha, const double *X, int incX,
                const double *Y, int incY, double *A, int lda);
double cblas_ddot(int N, const double *X, int incX, const double *Y, int incY);
void cblas_dtrsm(CBLAS_LAYOUT, CBLAS_SIDE, CBLAS_UPLO, CBLAS_TRANSPOSE, CBLAS_DIAG, int M, int N,
                 double alpha, const double *A, int lda, double *B, int ldb);
void cblas_ztrsm(CBLAS_LAYOUT, CBLAS_SIDE, CBLAS_UPLO, CBLAS_TRANSPOSE, CBLAS_DIAG, int M, int N,
                 const void *alpha, const void *A, int lda, void *B, int ldb);
void cblas_dsyrk(CBLAS_LAYOUT, CBLAS_UPLO, CBLAS_TRANSPOSE, int N, int K, double alpha,
                 const double *A, int lda, double beta, double *C, int ldc);
void cblas_dgemm_batch(CBLAS_LAYOUT, const CBLAS_TRANSPOSE *, const CBLAS_TRANSPOSE *,
                       const int *M, const int *N, const int *K, const double *alpha,
                       const double **A, const int *lda, const double **B, const int *ldb,
                       const double *beta, double **C, const int *ldc, int group_count,
                       const int *group_size);
void cblas_zgemm_batch(CBLAS_LAYOUT, const CBLAS_TRANSPOSE *, const CBLAS_TRANSPOSE *,
                       const int *M, const int *N, const int *K, const void *alpha,
                       const void **A, const int *lda, const void **B, const int *ldb,
                       const void *beta, void **C, const int *ldc, int group_count,
                       const int *group_size);

#ifdef __cplusplus
}
#endif
#endif
