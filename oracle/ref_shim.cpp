// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// In-process entry points into the UNMODIFIED reference decoder.  The reference
// translation unit (/root/reference/viterbi/viterbi_convolutional_code.cpp) is
// pulled in textually from where it lies (found through -I on the compile line,
// see oracle/Makefile); nothing from it is copied into this repository.  The
// result is oracle/_ref/libvcc_ref.so, used by tests/ to pin oracle/lva_oracle.cpp
// and by bench.py's cpu_baseline / --impl reference leg.
//
// The reference keeps its code parameters in file-scope globals (VCC:57-75) and
// is written as a one-shot process, so every call here re-establishes the
// globals the way a fresh process would see them.

#define main vcc_reference_main
#include "viterbi_convolutional_code.cpp"
#undef main

#include <cstring>

extern "C" {

// Returns 0 on success, 1 if set_conv_params rejected the parameters (VCC:264-415),
// 2 if the decoder threw (VCC:596-605).  out_bits holds list_size rows of msg_len
// bytes (0/1); *n_out rows are valid.
int vcc_ref_decode(const float *post, uint32_t nblk, uint32_t mem, uint32_t rate,
                   uint32_t msg_len, uint32_t list_size, uint32_t num_thr,
                   uint32_t max_deviation, int rc, const char *sync_marker_str,
                   uint32_t sync_period, uint8_t *out_bits, uint32_t *n_out) {
  rc_flag = (rc != 0);
  sync_marker_length = 0;  // fresh-process state
  sync_marker_period = 0;
  std::string marker = sync_marker_str ? sync_marker_str : "";
  int status = (marker == "")
                   ? set_conv_params((uint8_t)mem, (uint8_t)rate, msg_len)
                   : set_conv_params((uint8_t)mem, (uint8_t)rate, msg_len, marker,
                                     sync_period);
  if (status != 0) return 1;
  std::vector<crf_mat_t> pm(nblk);
  for (uint32_t t = 0; t < nblk; t++)
    for (int i = 0; i < 5; i++)
      for (int j = 0; j < 8; j++) pm[t][i][j] = post[(size_t)t * 40 + i * 8 + j];
  *n_out = 0;
  try {
    auto lists =
        decode_post_conv_parallel_LVA(pm, msg_len, list_size, num_thr, max_deviation);
    for (size_t k = 0; k < lists.size(); k++)
      for (uint32_t i = 0; i < msg_len; i++)
        out_bits[k * msg_len + i] = lists[k][i] ? 1 : 0;
    *n_out = (uint32_t)lists.size();
  } catch (const std::exception &) {
    return 2;
  }
  return 0;
}

// Convolutional encode of one message (VCC:450-499).  out_bits must hold
// 2*(msg_len+mem) bytes; *n_out receives the punctured length.
int vcc_ref_encode(const uint8_t *msg_bits, uint32_t msg_len, uint32_t mem,
                   uint32_t rate, uint8_t *out_bits, uint32_t *n_out) {
  rc_flag = false;
  sync_marker_length = 0;
  if (set_conv_params((uint8_t)mem, (uint8_t)rate, msg_len) != 0) return 1;
  std::vector<bool> msg(msg_len);
  for (uint32_t i = 0; i < msg_len; i++) msg[i] = msg_bits[i] != 0;
  try {
    std::vector<bool> enc = conv_encode(msg);
    for (size_t i = 0; i < enc.size(); i++) out_bits[i] = enc[i] ? 1 : 0;
    *n_out = (uint32_t)enc.size();
  } catch (const std::exception &) {
    return 2;
  }
  return 0;
}

// Code tables as the reference builds them (VCC:264-415), for table-builder parity.
int vcc_ref_tables(uint32_t mem, uint32_t rate, uint32_t msg_len, int rc,
                   uint32_t *G_out, uint32_t *init_final, uint8_t *punc,
                   uint32_t *punc_len, uint32_t *nstate_pos_out,
                   uint32_t *pos2msg /* 256 */) {
  rc_flag = (rc != 0);
  sync_marker_length = 0;
  if (set_conv_params((uint8_t)mem, (uint8_t)rate, msg_len) != 0) return 1;
  G_out[0] = G[0];
  G_out[1] = G[1];
  init_final[0] = initial_state_conv;
  init_final[1] = final_state_conv;
  *punc_len = puncturing_pattern_len;
  for (uint32_t i = 0; i < puncturing_pattern_len; i++) punc[i] = puncturing_pattern[i];
  *nstate_pos_out = nstate_pos;
  for (uint32_t i = 0; i < nstate_pos && i < 256; i++) pos2msg[i] = st_pos2msg_pos[i];
  return 0;
}

}  // extern "C"
