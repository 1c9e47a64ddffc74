/* TEST INFRASTRUCTURE ONLY -- CPU restatement of flappie's flip-flop transition posteriors
 * (/root/reference/flappie/src/decode.c:377-497 transpost_crf_flipflop with return_log = true,
 * logsumexpf util.h:188-190, log_row_normalise_inplace flappie_matrix.c:450-467), the step right
 * upstream of the list decoder.  Checker for lva_posteriors_batch; pinned
 * against the unmodified reference function (oracle/_ref/libflipflop_ref.so) in tests/test_flipflop.py.
 * Nothing under nanopore_dna_storage_b200/ may link or call this.
 *
 * trans / out: nblk x 40 floats, block-major: [to*8 + from] for the flip targets to = 0..3,
 * [32 + from] for the flop target of `from` (the .post layout, VCC:560-571). */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline float lse(float x, float y) { return fmaxf(x, y) + log1pf(expf(-fabsf(x - y))); }  /* util.h:188-190 */

int ffo_transpost(const float *trans, unsigned nblk, float *out) {
    enum { NB = 4, NS = 8 };
    float *fwd = calloc((size_t)(nblk + 1) * NS, sizeof(float));  /* make_flappie_matrix zero-fills: column 0 = 0 */
    if (!fwd) return 1;
    for (unsigned blk = 0; blk < nblk; blk++) {                   /* decode.c:398-424 */
        const float *t = trans + (size_t)blk * 40, *prev = fwd + (size_t)blk * NS;
        float *curr = fwd + (size_t)(blk + 1) * NS;
        for (int b2 = NB; b2 < NS; b2++) {
            curr[b2] = prev[b2] + t[32 + b2];
            curr[b2] = lse(curr[b2], prev[b2 - NB] + t[32 + b2 - NB]);
        }
        for (int b1 = 0; b1 < NB; b1++) {
            curr[b1] = t[b1 * NS] + prev[0];
            for (int from = 1; from < NS; from++) curr[b1] = lse(curr[b1], t[b1 * NS + from] + prev[from]);
        }
    }
    float mem[2 * NS] = {0};                                       /* decode.c:426-433 (calloc) */
    float *prev = mem, *curr = mem + NS;
    for (unsigned blk = nblk; blk > 0; blk--) {                    /* decode.c:436-483 */
        const float *t = trans + (size_t)(blk - 1) * 40, *f = fwd + (size_t)(blk - 1) * NS;
        float *o = out + (size_t)(blk - 1) * 40;
        { float *tmp = prev; prev = curr; curr = tmp; }
        for (int b1 = 0; b1 < NB; b1++)
            for (int st = 0; st < NS; st++) o[b1 * NS + st] = f[st] + prev[b1] + t[b1 * NS + st];
        for (int b = NB; b < NS; b++) {
            const int fb = b - NB;
            o[32 + b] = f[b] + prev[b] + t[32 + b];
            o[32 + fb] = f[fb] + prev[b] + t[32 + fb];
        }
        for (int b2 = NB; b2 < NS; b2++) {
            curr[b2] = prev[b2] + t[32 + b2];
            curr[b2 - NB] = prev[b2] + t[32 + b2 - NB];
        }
        for (int b1 = 0; b1 < NB; b1++)
            for (int from = 0; from < NS; from++) curr[from] = lse(curr[from], t[b1 * NS + from] + prev[b1]);
    }
    free(fwd);
    for (unsigned blk = 0; blk < nblk; blk++) {                    /* flappie_matrix.c:456-466 */
        float *o = out + (size_t)blk * 40;
        float s = o[0];
        for (int r = 1; r < 40; r++) s = lse(s, o[r]);
        for (int r = 0; r < 40; r++) o[r] -= s;
    }
    return 0;
}
