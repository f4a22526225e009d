/*
 * oracle/compact_ref.c — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Plain-C restatement of the integer/byte part of the reference path and a naive fp32 attention over an index
 * list, used by tests/ to check the CUDA compaction bit-exactly and as an independent cross-check of the
 * torch-based port in oracle/reference_port.py.  Reference lines are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* StoryDiffusion/utils/gradio_utils.py:260-278 — distinct mask row i of cal_attn_mask_xl:
 * the sampled vector restricted to columns < id_length*n, own block [i*n, (i+1)*n) forced True. */
void csa_ref_frame_row(const uint8_t* sample, int total_length, int id_length, int n, int i, uint8_t* row_out) {
  const int cols = total_length * n;
  for (int j = 0; j < cols; ++j) {
    uint8_t v = sample[j] != 0;
    if (j >= id_length * n) v = 0;
    if (j >= i * n && j < (i + 1) * n) v = 1;
    row_out[j] = v;
  }
}

/* torch.nonzero(row) — ascending attended columns; returns the count.  This is what the mask slicing of
 * StoryDiffusion/Comic_Generation.py:105-114 selects for every query of the frame. */
int csa_ref_nonzero(const uint8_t* row, int n_cols, int32_t* idx_out) {
  int c = 0;
  for (int j = 0; j < n_cols; ++j)
    if (row[j]) idx_out[c++] = j;
  return c;
}

/* Dense-mask form of the block-row premise: 1 if every row of the (rows x cols) mask equals the first row of
 * its block of block_n rows (gradio_utils.py:285-286 builds the mask that way), else 0. */
int csa_ref_blocks_uniform(const uint8_t* mask, int64_t row_stride, int rows, int cols, int block_n) {
  for (int r = 0; r < rows; ++r) {
    const uint8_t* ref = mask + (int64_t)(r / block_n * block_n) * row_stride;
    const uint8_t* cur = mask + (int64_t)r * row_stride;
    for (int j = 0; j < cols; ++j)
      if ((ref[j] != 0) != (cur[j] != 0)) return 0;
  }
  return 1;
}

/* softmax(q k^T * scale) v for ONE head over the keys listed in idx (StoryDiffusion/Comic_Generation.py:175-177
 * restricted to the attended columns).  q: (nq, d) row stride q_ld; k, v: row stride kv_ld; out: (nq, d) row
 * stride o_ld.  Double accumulation; meant for small cases only. */
void csa_ref_attention_f32(const float* q, int64_t q_ld, const float* k, const float* v, int64_t kv_ld,
                           const int32_t* idx, int n_keys, int nq, int d, float scale, float* out, int64_t o_ld) {
  double* s = (double*)malloc(sizeof(double) * (size_t)n_keys);
  for (int i = 0; i < nq; ++i) {
    double mx = -INFINITY;
    for (int t = 0; t < n_keys; ++t) {
      const float* kr = k + (int64_t)idx[t] * kv_ld;
      double a = 0.0;
      for (int e = 0; e < d; ++e) a += (double)q[i * q_ld + e] * (double)kr[e];
      s[t] = a * scale;
      if (s[t] > mx) mx = s[t];
    }
    double l = 0.0;
    for (int t = 0; t < n_keys; ++t) {
      s[t] = exp(s[t] - mx);
      l += s[t];
    }
    for (int e = 0; e < d; ++e) {
      double a = 0.0;
      for (int t = 0; t < n_keys; ++t) a += s[t] * (double)v[(int64_t)idx[t] * kv_ld + e];
      out[i * o_ld + e] = (float)(a / l);
    }
  }
  free(s);
}
