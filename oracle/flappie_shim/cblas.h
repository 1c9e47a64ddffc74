/* TEST INFRASTRUCTURE: stand-in for <cblas.h>, which this image does not have.  The only reference
 * functions the oracle calls (transpost_crf_flipflop and the matrix helpers it uses) never reach
 * BLAS; the two routines flappie_matrix.c mentions are declared here and defined in
 * flipflop_ref_shim.c to abort if anything ever calls them. */
#ifndef ORACLE_CBLAS_STUB_H
#define ORACLE_CBLAS_STUB_H
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
void cblas_sgemm(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE ta, const enum CBLAS_TRANSPOSE tb,
                 const int m, const int n, const int k, const float alpha, const float *a, const int lda,
                 const float *b, const int ldb, const float beta, float *c, const int ldc);
void cblas_sgemv(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE ta, const int m, const int n,
                 const float alpha, const float *a, const int lda, const float *x, const int incx,
                 const float beta, float *y, const int incy);
#endif
