/* TEST INFRASTRUCTURE: C entry point around the UNMODIFIED reference transpost_crf_flipflop
 * (/root/reference/flappie/src/decode.c:377-497), compiled from the sources where they lie
 * (oracle/Makefile, target _ref/libflipflop_ref.so).  Nothing of the reference is copied. */
#include <stdlib.h>
#include <string.h>
#include "decode.h"
#include "flappie_matrix.h"
#include "cblas.h"

void cblas_sgemm(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE ta, const enum CBLAS_TRANSPOSE tb,
                 const int m, const int n, const int k, const float alpha, const float *a, const int lda,
                 const float *b, const int ldb, const float beta, float *c, const int ldc) { abort(); }
void cblas_sgemv(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE ta, const int m, const int n,
                 const float alpha, const float *a, const int lda, const float *x, const int incx,
                 const float beta, float *y, const int incy) { abort(); }

/* layers.c (the network layers, BLAS all over) is not compiled; transpost_crf_flipflop references this
 * one function of it, on the return_log == false branch the oracle never takes */
void exp_activation_inplace(flappie_matrix C) { (void)C; abort(); }
/* same for a helper of layers.c that another decoder in decode.c (not the posterior) calls */
size_t nbase_from_flipflop_nparam(size_t nparam) { (void)nparam; abort(); }

/* trans: nblk x 40 floats (block-major, the .post / A1 order); out: nblk x 40 log-posteriors */
int ffr_transpost(const float *trans, unsigned nblk, float *out) {
    flappie_matrix t = make_flappie_matrix(40, nblk);
    if (!t) return 1;
    for (unsigned b = 0; b < nblk; b++) memcpy(t->data.f + (size_t)b * t->stride, trans + (size_t)b * 40, 160);
    flappie_matrix p = transpost_crf_flipflop(t, true);
    if (!p) { free_flappie_matrix(t); return 2; }
    for (unsigned b = 0; b < nblk; b++) memcpy(out + (size_t)b * 40, p->data.f + (size_t)b * p->stride, 160);
    free_flappie_matrix(p);
    free_flappie_matrix(t);
    return 0;
}
