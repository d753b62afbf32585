// Launchers of the CUDA kernels (C++ linkage, internal to the library).
//
// Device layout of a batch: reads are concatenated.  Read r owns samples
// [samp_off[r], samp_off[r] + nsample[r]) of `raw` and columns ("blocks")
// [col_off[r], col_off[r] + nblock[r]) of every activation matrix.  Activations are
// column-major like the reference's scrappie_matrix: one column = one time step, the
// features of a column are contiguous ([total_cols][nfeature], no padding lanes except
// in the posterior, which keeps the reference's stride = 4 * ceil(nstate / 4)).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sb2_internal.h"

namespace sb2 {

struct BatchDims {
    int nread;
    int total_cols;
    int max_cols;                   // longest read of the batch, in columns
    const int *nsample;             // device [nread]
    const int *nblock;              // device [nread]
    const int64_t *samp_off;        // device [nread]
    const int *col_off;             // device [nread + 1]
    const int *col_read;            // device [total_cols]: owning read of each column (may be null)
};

// conv + activation (src/layers.c:159-246 + :60-68 / :15-24)
void launch_conv_act(const float *raw, const BatchDims &d, const sb2_conv_tail *tails,
                     const float *taps /*[winlen][nf]*/, const float *bias, int winlen, int nf,
                     int stride, int act, float *out, cudaStream_t s);

// C[col][m] = f((base + sum_k W[m][k] * (X[col][k] / xdiv)) / cdiv); base = b[m], or the old C[col][m]
// when `accumulate` (second product of affine_map2); f (act) = 0 identity, 1 exp, 2 tanh
// (affine_map / affine_map2 src/scrappie_matrix.c:323-383; softmax_with_temperature src/layers.c:340-357;
//  feedforward2_tanh src/layers.c:359-371)
void launch_affine(const float *X, int ncol, int K, const float *W, int ldw, const float *b, int M,
                   float *C, int ldc, float xdiv, float cdiv, int act, int accumulate, cudaStream_t s);

// row_normalise_inplace + robustlog_activation_inplace (src/scrappie_matrix.c:385-407,
// src/layers.c:79-94) over the exp'd head output, padding lanes included
void launch_softmax_finish(float *post, int ncol, int nstate, int ostride, float min_prob,
                           int return_log, cudaStream_t s);

// gru_forward / gru_backward (src/layers.c:373-527), fp32 FFMA path
void launch_gru_scan_ffma(const float *Xin, const float *sW, const float *sW2, const float *resid,
                          float *out, const BatchDims &d, int H, int backward, cudaStream_t s);

// globalnorm (src/layers.c:835-889): trans -= logZ / T per read
void launch_globalnorm(float *trans, const BatchDims &d, int ostride, cudaStream_t s);

// decode_transducer + viterbi_local_backtrace (src/decode.c:58-98, :123-365)
void launch_decode_transducer(const float *post, const BatchDims &d, int nstate, int ostride,
                              float stay_pen, float skip_pen, float local_pen, int allow_slip,
                              uint8_t *tb, int *tb_end, int *path, float *score, cudaStream_t s);

// the same for 1024 histories without slip, one warp per read (kernels_decode.cu)
void launch_decode_transducer_warp(const float *post, const BatchDims &d, int ostride, float stay_pen,
                                   float skip_pen, float local_pen, uint8_t *tb, int *tb_end, int *path,
                                   float *score, cudaStream_t s);

// posterior_crf (src/decode.c:928-1012): post = (total_cols + nread) columns of 8 floats, read r starts at
// column col_off[r] + r and owns nblock[r] + 1 columns
void launch_posterior_crf(const float *trans, const BatchDims &d, int ostride, float *post, cudaStream_t s);

// signal preparation (kernels_prep.cu): trim_and_segment_raw (src/scrappie_common.c:5-73) -> start_end[2 r], [2 r + 1]
// (both 0 when nothing is left), and medmad_normalise_array (src/util.c:190-204) from src + src_off[r] to dst + dst_off[r]
void launch_trim(const float *raw, const int64_t *off, const int *nsample, int nread, int chunk, float perc,
                 int trim_start, int trim_end, float *mads, const int64_t *mads_off, int *start_end, cudaStream_t s);
void launch_medmad(const float *src, const int64_t *src_off, float *dst, const int64_t *dst_off, const int *nsample,
                   int nread, cudaStream_t s);

// events model (kernels_lstm.cu): window(features, 3, 1) (src/layers.c:119-146) over [col][4] features -> [col][12],
// and lstm_forward / lstm_backward (src/layers.c:673-832): Xin [col][4H], sW [4H][H], peep [3H] -> out [col][H]
void launch_window3(const float *feat, float *out, const BatchDims &d, cudaStream_t s);
int launch_lstm_scan(const float *Xin, const float *sW, const float *peep, float *out, const BatchDims &d, int H,
                     int backward, cudaStream_t s);

// map_to_sequence_{viterbi,forward}[_banded] (src/decode.c:1420-1964), kernels_map.cu.  buf: 2 * (seqlen + 2)
// floats; tb (seqlen bytes per block) / tb_end (1 byte per block) / path only for the unbanded Viterbi path
void launch_map_to_sequence(const float *lp, int nblock, int nst, int stride, float stay_pen, float skip_pen,
                            float local_pen, const int *seq, int seqlen, const int *low, const int *high, int forward,
                            float *buf, uint8_t *tb, uint8_t *tb_end, float *score, int *path, cudaStream_t s);

// decode_crf (src/decode.c:836-893)
void launch_decode_crf(const float *trans, const BatchDims &d, int ostride, uint8_t *tb, int *path,
                       float *score, cudaStream_t s);

// homopolymer_path + overlapper (or crfpath_to_basecall) per read on the device; bases[r * bases_stride ...] receives a
// NUL-terminated string, nbase[r] its length (-1 when the path holds no k-mer)
void launch_finish_reads(const float *post, const BatchDims &d, int nstate, int ostride, int head, int homopolymer,
                         int klen, const int *path_in, int *path_work, char *bases, int bases_stride, int *nbase,
                         cudaStream_t s);

// small-M affine map for the CRF head (M <= 32 rows): C[col][0:M] = b + W^T X[col], lanes M..ldc-1 zeroed
void launch_small_head(const float *X, int ncol, int K, const float *W, const float *b, int M, float *C, int ldc,
                       cudaStream_t s);

// gather post[col][state] pairs (homopolymer fix-up needs a few posterior entries)
void launch_gather(const float *post, int ostride, const int *col_state_pairs, int n, float *out, cudaStream_t s);

// ---- tensor-core path (kernels_tc.cu) ----
// gru_forward / gru_backward (+ residual) on tcgen05 with the recurrent weights resident in TMEM.
// math: 0 cephes gates, 5 SFU ex2 + Newton-refined reciprocal.  Batches of >= 48 reads run eight reads per group,
// smaller ones four (scan_reads_per_group).  xgrp != nullptr: Xin is in SCAN ORDER -- read r, step s (s = t forward,
// T - 1 - t backward) at row xgrp[r / RPG] + s * RPG + r % RPG -- and every group-step is one TMA copy; nullptr: Xin is
// read-major like every other activation matrix.
int scan_reads_per_group(int nread);
int scan_launch_priority();
int launch_gru_scan_tc(const float *Xin, const long long *xgrp, const float *sW, const float *sW2, const float *resid,
                       float *out, const BatchDims &d, int H, int backward, int math, long long *trace, cudaStream_t s);
// src_col tables of the scan-ordered Xin: src_f[row] / src_b[row] = input column of every Xin row for forward / backward
// layers (-1 for the unused rows of ragged groups); nrow = rows of the layout
int launch_scan_rows(const BatchDims &d, const long long *xgrp, int rpg, size_t nrow, int *src_f, int *src_b, cudaStream_t s);
// per-device kernel attributes (dynamic shared-memory limits); called by sb2_engine_create for its device
int configure_scan_kernels();
int configure_gemm_kernels();
int configure_v1_kernels();
// D[128][N] = A[128][K] B[N][K]^T through the scan's operand path (validation / latency probe)
int launch_tc_selftest(const float *A, const float *B, float *D, int K, int N, int reps, long long *cycles,
                       cudaStream_t s);

// ---- tensor-core affine maps (kernels_gemm.cu) ----
// weight image: ntile tiles of `rows` output units, hi + lo fp16, canonical UMMA layout
size_t gemm_image_bytes(int ntile, int rows, int K);
void build_gemm_image(const float *W, int ldw, int M, int K, int rows, int ntile, uint8_t *img);
// feedforward_linear for a GRU layer: C[col][0:3H] = b + iW^T X[col]  (src/layers.c:248-252)
// src_col (may be null): the kernel makes `ncol` OUTPUT rows; row i is computed from input column src_col[i] (-1: the
// row is not used) -- the scan-ordered Xin layout; null = row i from column i
int launch_affine_tc(const float *X, int ncol, int H, const uint8_t *wimg, const float *bias, float *C, const int *src_col,
                     cudaStream_t s);

// fused output head for 1025-state models: FF GEMM -> softmax with temperature -> robust log
// (src/layers.c:340-357, :79-94); wimg = build_gemm_image(FF_W, K, 1024, K, 128, 8, .), w_stay = FF_W row 1024
size_t head_image_bytes(int K);
int launch_head_softmax_tc(const float *X, int ncol, int K, const uint8_t *wimg, const float *w_stay, const float *bias,
                           float *post, int ostride, float xdiv, float cdiv, float min_prob, int return_log,
                           int exact_math, cudaStream_t s);

// overwrite a buffer larger than L2 (benchmark hygiene)
void launch_flush(float *buf, size_t nfloat, cudaStream_t s);

}  // namespace sb2
