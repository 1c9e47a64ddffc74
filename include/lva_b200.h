/* lva_b200.h -- C ABI of the B200 list-Viterbi decoder (liblva_b200.so).
 *
 * The reference (shubhamchandak94/nanopore_dna_storage) has no library or FFI boundary for this
 * path: its decoder is one function inside one executable,
 *     decode_post_conv_parallel_LVA   viterbi/viterbi_convolutional_code.cpp:126-129, 589-858
 * configured by file-scope globals that set_conv_params fills in (:57-75, :264-415) and fed by
 * read_crf_post (:553-575).  This header is the cut a maintainer would bind if that function
 * were lifted out of `main` (:137-262): every entry point below names the reference code it
 * replaces.  Plain C types only; caller-owned buffers; int status codes; one handle per GPU;
 * a handle may be used by one host thread at a time.
 *
 * There is no CPU fallback: lva_create fails if no CUDA device is usable.
 */
#ifndef LVA_B200_H_
#define LVA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LVA_DEFAULT_DEVIATION 0xFFFFFFFFu /* "--max-deviation" absent: msg_len+mem_conv+1 (VCC:238-242) */

/* status codes (lva_create / per call) */
enum {
  LVA_OK = 0,
  LVA_ERR_MEM_CONV = 1,      /* VCC:290-292 "Invalid mem_conv (allowed: 6, 8, 11, 14)" */
  LVA_ERR_RATE = 2,          /* VCC:336-338 "Invalid rate parameter ..." */
  LVA_ERR_ODD_LENGTH = 3,    /* VCC:353-357 "Output length not even. ..." */
  LVA_ERR_MARKER_LONG = 4,   /* VCC:391-394 */
  LVA_ERR_MARKER_PERIOD = 5, /* VCC:396-399 */
  LVA_ERR_MARKER_CHAR = 6,   /* VCC:409-411 */
  LVA_ERR_TOO_MANY_STATES = 20, /* VCC:595-597 */
  LVA_ERR_POST_TOO_SMALL = 21,  /* VCC:600-601  per-read status: T < nstate_pos + 1 */
  LVA_ERR_MSG_TOO_LONG = 22,    /* VCC:604-605 (and msg_len + mem_conv must fit the 256-bit register) */
  LVA_ERR_LIST_SIZE = 23,       /* list_size 0 or above LVA_MAX_LIST */
  LVA_ERR_BAD_ARG = 30,
  LVA_ERR_NO_DEVICE = 40,       /* no usable CUDA device: there is no CPU path */
  LVA_ERR_CUDA = 41,
  LVA_ERR_OOM = 42,
  LVA_ERR_CODE_STRUCTURE = 43   /* generator polynomials violate the one-predecessor-per-base structure */
};

#define LVA_MAX_LIST 64u

/* The reference CLI's decode parameters (cxxopts block, VCC:138-172) + globals (VCC:57-75). */
typedef struct lva_params {
  uint32_t mem_conv;      /* --mem-conv  6 | 8 | 11 | 14 */
  uint32_t rate;          /* -r          1,2,3,4,5,7 = 1/2,2/3,3/4,4/5,5/6,7/8 */
  uint32_t msg_len;       /* --msg-len */
  uint32_t list_size;     /* -l */
  uint32_t max_deviation; /* --max-deviation, or LVA_DEFAULT_DEVIATION */
  int32_t rc;             /* --rc */
  const char *sync_marker;/* --sync-marker ("" or NULL for none) */
  uint32_t sync_period;   /* --sync-period */
  int32_t device;         /* CUDA device ordinal */
  uint32_t max_slots;     /* reads decoded concurrently (0 = choose from free memory) */
  uint32_t flags;         /* LVA_FLAG_* */
} lva_params;

#define LVA_FLAG_EXACT_ONLY 1u /* run every state through the exact heap-order merge (validation mode) */
/* Band start of VCC:677-679 with the product t/nblk*nstate_pos rounded BEFORE max_deviation is
 * subtracted.  Default (flag clear) is the fused multiply-subtract g++ -O3 -march=native emits for the
 * reference on FMA hardware (install.sh:10); set this to match a reference built with
 * -ffp-contract=off, with clang, or for a CPU without FMA.  The two differ by one trellis row on a few
 * steps of some read lengths (e.g. C2 reads whose length is a multiple of 29). */
#define LVA_FLAG_UNFUSED_BAND 2u

typedef struct lva_handle lva_handle;

/* Replaces set_conv_params (VCC:264-415) + the table builds at the top of
 * decode_post_conv_parallel_LVA (VCC:607-650).  Returns NULL on failure and writes an LVA_ERR_*
 * code to *status (may be NULL) and the reference's message text to err (may be NULL). */
lva_handle *lva_create(const lva_params *params, int *status, char *err, size_t err_len);
void lva_destroy(lva_handle *h);
const char *lva_last_error(const lva_handle *h);

/* Replaces read_crf_post + decode_post_conv_parallel_LVA (VCC:553-575, 589-858) for a batch of
 * reads.  post = concatenated posterior blocks, 40 floats each in the file order
 * (VCC:560-571: to*8+from for flip targets, 32+from for the flop target); read i owns blocks
 * [block_offsets[i], block_offsets[i+1]).  Outputs (caller-owned):
 *   out_bits   [n_reads][list_size][msg_len]  0/1 bytes, list order and bit order of VCC:826-836
 *   out_scores [n_reads][list_size]           path metric of each entry (reference: LVA_path_t.score)
 *   out_counts [n_reads]                      entries produced (<= list_size; 0 = empty list)
 *   out_status [n_reads]                      LVA_OK or LVA_ERR_POST_TOO_SMALL per read
 * Any of out_scores / out_status may be NULL.  `post` is HOST memory here. */
int lva_decode_batch(lva_handle *h, const float *post, const uint64_t *block_offsets,
                     uint32_t n_reads, uint8_t *out_bits, float *out_scores,
                     uint32_t *out_counts, int32_t *out_status);

/* Same, with `post` already resident in device memory of the handle's GPU (for callers that
 * produce posteriors on the GPU, and for kernel-only timing).  Offsets/outputs stay on the host. */
int lva_decode_batch_device(lva_handle *h, const float *post_dev, const uint64_t *block_offsets,
                            uint32_t n_reads, uint8_t *out_bits, float *out_scores,
                            uint32_t *out_counts, int32_t *out_status);

/* Same as lva_decode_batch / lva_decode_batch_device for reads that are arbitrary block ranges of one
 * posterior array: read i owns blocks [block_starts[i], block_ends[i]) (ranges may leave gaps or
 * overlap).  This is what helper.truncate_post_file (helper.py:211-224) does with a file copy: pass
 * the [start, end + 1) ranges lva_trim_batch found and the untrimmed posteriors stay where they are
 * (post_on_device != 0: `post` is device memory, e.g. the output of lva_posteriors_batch_device). */
int lva_decode_ranges(lva_handle *h, const float *post, int32_t post_on_device, const uint64_t *block_starts,
                      const uint64_t *block_ends, uint32_t n_reads, uint8_t *out_bits, float *out_scores,
                      uint32_t *out_counts, int32_t *out_status);

/* lva_decode_batch plus the check the reference's drivers run on every list afterwards
 * (helper.decode_list_CRC_index, helper.py:371-388; framing helper.py:257-264), computed on the GPU
 * for all candidates: strip the pad bit (pad != 0), left-pad to whole bytes, CRC-8 (poly 0x07, init 0)
 * of all but the last byte must equal the last byte; index = PRP^-1 of the low 12 bits of the first
 * two bytes.  out_crc_ok / out_index are [n_reads][list_size] in list order; out_choice[i] is the
 * first entry with a matching CRC and index < num_oligos, or -1 (any of the three may be NULL).
 * The list outputs are unchanged: the CRC parameters are not options of the reference tool. */
int lva_decode_batch_crc(lva_handle *h, const float *post, const uint64_t *block_offsets, uint32_t n_reads,
                         int32_t pad, uint32_t num_oligos, uint8_t *out_bits, float *out_scores,
                         uint32_t *out_counts, int32_t *out_status, uint8_t *out_crc_ok,
                         uint16_t *out_index, int32_t *out_choice);

/* conv_encode (VCC:450-499): msg = msg_len 0/1 bytes; out gets 2 bits per base.  Host only. */
int lva_encode(uint32_t mem_conv, uint32_t rate, const uint8_t *msg, uint32_t msg_len,
               uint8_t *out_bits, uint32_t out_cap, uint32_t *n_out);

/* set_conv_params' acceptance test alone (VCC:264-357): LVA_OK or LVA_ERR_MEM_CONV / _RATE / _ODD_LENGTH.
 * Host only. */
int lva_check_params(uint32_t mem_conv, uint32_t rate, uint32_t msg_len);

/* The code tables set_conv_params builds (VCC:57-75, 264-386), for callers and parity tests.
 * Host only.  punc must hold 16 bytes, pos2msg 260 words.  Returns like lva_check_params. */
typedef struct lva_code_tables {
  uint32_t G[2], initial_state, final_state, punc_len, nstate_pos;
  uint8_t punc[16];
  uint32_t pos2msg[260];   /* st_pos2msg_pos */
  uint32_t valid_mask[260], valid_value[260]; /* is_valid_state (VCC:944-978) as (c & mask) == value per row */
  uint8_t history_words[260];                 /* 32-bit message-history words carried per list entry per row */
  uint8_t first_run_row[260];                 /* 1: some message prefix reaches this row with all its bases equal (row 0 included):
                                                 only there can one message sit in two crf states of a conv state */
} lva_code_tables;
int lva_get_code_tables(uint32_t mem_conv, uint32_t rate, uint32_t msg_len, int32_t rc,
                        const char *sync_marker, uint32_t sync_period, lva_code_tables *out);

/* band of VCC:677-679 for step t of a read of nblk blocks: [*start, *end) (fused form; lva_band_ex
 * takes LVA_FLAG_UNFUSED_BAND in flags) */
void lva_band(uint32_t t, uint32_t nblk, uint32_t nstate_pos, uint32_t max_deviation,
              uint32_t *start, uint32_t *end);
void lva_band_ex(uint32_t t, uint32_t nblk, uint32_t nstate_pos, uint32_t max_deviation, uint32_t flags,
                 uint32_t *start, uint32_t *end);

/* Geometry a caller may want (nstate_pos of VCC:74; bases per oligo = nstate_pos - 1). */
uint32_t lva_nstate_pos(const lva_handle *h);
uint32_t lva_slots(const lva_handle *h);

/* Work in the reference's unit for a read of T blocks. */
typedef struct lva_work {
  double state_updates; /* (step, in-band row, reachable conv state, crf state) */
  double edge_acs;      /* state updates x incoming edges */
  double state_bytes;   /* algorithmic state traffic: one read + one write of each live list */
} lva_work;
int lva_count_work(const lva_handle *h, uint32_t T, lva_work *out);

/* Timing of the most recent decode call, measured with CUDA events on the handle's stream. */
typedef struct lva_timing {
  float total_ms;   /* whole call on the device stream (H2D + kernels + D2H) */
  float kernel_ms;  /* trellis step launches only */
  float h2d_ms, d2h_ms;
  uint64_t launches;        /* kernel launches issued */
  uint64_t step_launches;   /* of which trellis step kernels */
  uint64_t h2d_bytes, d2h_bytes;
  uint64_t exact_fallbacks; /* states resolved by the exact heap-order path after a score tie */
  uint64_t chain_inserts;   /* distinct (chain id, chunk) keys interned by the call (all reads) */
  uint64_t chain_keys_max;  /* ... the most any one read needed */
  uint64_t chain_table_slots; /* slots of one read's chain table */
  uint64_t table_regrows;   /* times the call was repeated with larger chain tables */
  uint64_t structure_check_failures; /* developer check (LVA_DBG bit 3) of the two-stage merge's precondition; 0 */
} lva_timing;
int lva_get_timing(const lva_handle *h, lva_timing *out);

/* ---- posterior trimming -------------------------------------------------
 * Replaces, for a batch of reads, helper.find_barcode_pos_in_post (helper.py:157-209) called for the
 * forward barcodes and for their reverse complements, and the orientation / size rule of the driver
 * (generate_decoded_lists.py:68-83).  helper.truncate_post_file (helper.py:211-224) needs no
 * counterpart: pass [out_start, out_end + 1) as the read's block range to lva_decode_ranges.
 *   basecalls / base_offsets   concatenated basecalls (ACGTN bytes), read i = [base_offsets[i], base_offsets[i+1])
 *   trans / trans_offsets      concatenated flappie .trans tables (flappie.c:278-282: one block index per base)
 *   start_barcode, end_barcode as given to the driver; the RC pair is rc(end), rc(start) (generate_decoded_lists.py:33-34)
 *   min_blocks                 mem_conv + msg_len + 1 (generate_decoded_lists.py:76)
 * Outputs per read (host, caller-owned): out_start / out_end = first / last payload block (inclusive;
 * -1 on failure), out_rc = 1 when the reverse-complement barcodes match better, out_status = LVA_TRIM_*,
 * out_dist (may be NULL) = [n_reads][4] edit distances start/end forward, start/end rc (0xFFFFFFFF = the
 * reference's inf).  Returns LVA_OK or LVA_ERR_BAD_ARG / _NO_DEVICE / _OOM / _CUDA.  No CPU path. */
#define LVA_TRIM_MAX_BARCODE 64
enum {
  LVA_TRIM_OK = 0,
  LVA_TRIM_TOO_SHORT = 1,      /* helper.py:177-179 "Too short read" */
  LVA_TRIM_BARCODE_ORDER = 2,  /* helper.py:207-209 "Barcode removal failure" (end before start) */
  LVA_TRIM_TOO_FEW_BLOCKS = 3, /* generate_decoded_lists.py:76 "Failure in barcode removing." */
  LVA_TRIM_NO_WINDOW = 4       /* no window to search / .trans shorter than the basecall: the reference raises */
};
int lva_trim_batch(int32_t device, const char *basecalls, const uint64_t *base_offsets,
                   const uint32_t *trans, const uint64_t *trans_offsets, uint32_t n_reads,
                   const char *start_barcode, const char *end_barcode, uint32_t min_blocks,
                   int32_t *out_start, int32_t *out_end, int32_t *out_rc, int32_t *out_status,
                   uint32_t *out_dist);

/* ---- flip-flop transition posteriors ------------------------------------
 * Replaces transpost_crf_flipflop(trans, return_log = true) (flappie/src/decode.c:377-497, with
 * logsumexpf util.h:188-190 and log_row_normalise_inplace flappie_matrix.c:450-467) for a batch of
 * reads: from the CRF transition weights of the flappie network to the log-posteriors the decoder
 * reads.  trans and out_post: concatenated blocks of 40 floats in the .post order (flappie's
 * column-major 40 x nblk matrix is exactly that), read i = blocks [block_offsets[i], block_offsets[i+1]).
 * Sums are folded in the reference's order; results differ from the CPU only by the last-bit
 * behaviour of expf / log1pf.  The _device variant takes and fills device memory of `device`, so
 * the posteriors can go straight into lva_decode_batch_device.  Returns LVA_OK or an LVA_ERR_*. */
int lva_posteriors_batch(int32_t device, const float *trans, const uint64_t *block_offsets, uint32_t n_reads,
                         float *out_post);
int lva_posteriors_batch_device(int32_t device, const float *trans_dev, const uint64_t *block_offsets,
                                uint32_t n_reads, float *out_post_dev);

/* CUDA devices visible to the process (0 when there is none: there is no CPU path). */
int lva_device_count(void);

const char *lva_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LVA_B200_H_ */
