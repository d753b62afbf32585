/* libscrappie_b200 -- B200-native drop-in for the `scrappie raw` hot path.
 *
 * C-ABI of libscrappie_b200.so.  Part 1 re-exports, with identical names, argument
 * meaning, ownership and error behaviour, the libscrappie symbols that sit on the raw
 * basecalling path (reference: interface/scrappie.h:47-52, python/pyscrap.h:1-27,62,
 * src/networks.h:22-49, src/decode.h:13-30, src/homopolymer.h:13-14,
 * src/scrappie_matrix.h:31-41, src/scrappie_common.h, src/util.h:262).  Part 2 is the
 * batch interface the reference does not have: many reads per call, which is what the
 * GPU needs; the single-read symbols of part 1 are batches of one.
 *
 * The network forward pass and both Viterbi decoders run as sm_100a CUDA kernels.
 * Containers, signal trimming / normalisation and the O(T) integer post-processing
 * (overlapper, crfpath_to_basecall, homopolymer_path) stay on the host, as in the
 * reference.  There is no CPU fallback for the GPU stages: without a usable CUDA
 * device the posterior / decode entry points fail (NULL / NAN) and say so on stderr.
 */
#ifndef SCRAPPIE_B200_H
#define SCRAPPIE_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ===================================================================== */
/* Part 1 -- libscrappie-compatible surface                               */
/* ===================================================================== */

/* src/scrappie_structures.h:24-30.  Passed BY VALUE, borrowed, never modified by
 * the posterior functions. */
typedef struct {
    char *uuid;
    size_t n;
    size_t start;
    size_t end;
    float *raw;
} raw_table;

/* src/scrappie_structures.h:8-22, interface/scrappie.h:13-27 */
typedef struct {
    uint64_t start;
    float length;
    float mean;
    float stdv;
    int pos;
    int state;
} event_t;

typedef struct {
    size_t n;
    size_t start;
    size_t end;
    event_t *event;
} event_table;

/* src/scrappie_matrix.h:10-16.  Column-major fp32, every column padded to a multiple
 * of 4 floats (stride = 4 * nrq), 16-byte aligned, zero initialised.  The reference's
 * union {__m128 *v; float *f;} is a single pointer; `f` here has the same offset. */
typedef struct {
    size_t nr, nrq, nc, stride;
    union {
        void *v;
        float *f;
    } data;
} _Mat;
typedef _Mat *scrappie_matrix;
typedef _Mat const *const_scrappie_matrix;

/* src/networks.h:8-14 */
enum raw_model_type {
    SCRAPPIE_MODEL_RAW = 0,
    SCRAPPIE_MODEL_RGRGR_R9_4,
    SCRAPPIE_MODEL_RGRGR_R9_4_1,
    SCRAPPIE_MODEL_RGRGR_R10,
    SCRAPPIE_MODEL_RNNRF_R9_4,
    SCRAPPIE_MODEL_INVALID
};

/* src/homopolymer.h:6-10 */
enum homopolymer_calculation {
    HOMOPOLYMER_NOCHANGE = 0,
    HOMOPOLYMER_MEAN,
    HOMOPOLYMER_INVALID
};

/* -- host containers: src/scrappie_matrix.c:11-136 -- */
scrappie_matrix make_scrappie_matrix(size_t nr, size_t nc);
scrappie_matrix remake_scrappie_matrix(scrappie_matrix M, size_t nr, size_t nc);
scrappie_matrix copy_scrappie_matrix(const_scrappie_matrix M);
scrappie_matrix free_scrappie_matrix(scrappie_matrix mat);      /* always returns NULL */
void zero_scrappie_matrix(scrappie_matrix M);
scrappie_matrix mat_from_array(const float *x, size_t nr, size_t nc);
float *array_from_scrappie_matrix(const_scrappie_matrix mat);

/* -- signal preparation (host): src/scrappie_common.c:5-73, src/util.c:92-204 -- */
void medmad_normalise_array(float *x, size_t n);
float medianf(const float *x, size_t n);
float madf(const float *x, size_t n, const float *med);
void quantilef(const float *x, size_t nx, float *p, size_t np);
/* frees rt.raw and returns a zeroed table when the trimmed range is empty */
raw_table trim_and_segment_raw(raw_table rt, size_t trim_start, size_t trim_end,
                               size_t varseg_chunk, float varseg_thresh);
raw_table trim_raw_by_mad(raw_table rt, size_t chunk_size, float perc);

/* -- model registry: src/networks.c:17-127, python/build.py:34-44 -- */
typedef scrappie_matrix (*posterior_function_ptr)(const raw_table, float, float, float, bool);
enum raw_model_type get_raw_model(const char *modelstr);
const char *raw_model_string(const enum raw_model_type model);
int get_raw_model_stride(const enum raw_model_type model);
int get_raw_model_stride_from_string(const char *modelstr);
posterior_function_ptr get_posterior_function(const enum raw_model_type model);

/* -- network forward (GPU): src/networks.c:250-394, :567-615.
 *    Returns a new matrix [nstate x nblock] owned by the caller, or NULL. -- */
/* raw_r94 (interface/scrappie.h:49-51, src/networks.c:196-247): two bidirectional GRU pairs */
/* src/event_detection.h:6-25, src/event_detection.c:270-320: segmentation of a raw signal into events (host code, as in
 * the reference); the table's `event` array is calloc'd and owned by the caller */
typedef struct {
    size_t window_length1;
    size_t window_length2;
    float threshold1;
    float threshold2;
    float peak_height;
} detector_param;
event_table detect_events(raw_table const rt, detector_param const edparam);

/* interface/scrappie.h:47-48, src/networks.c:146-194: the events (LSTM) model.  Features are made on the host
 * (nanonet_features_from_events, src/nnfeatures.c:76-115, exported as well); window, LSTM layers and head on the GPU */
scrappie_matrix nanonet_features_from_events(const event_table et, bool normalise);
scrappie_matrix nanonet_posterior(const event_table events, float min_prob,
                                  float tempW, float tempb, bool return_log);
scrappie_matrix nanonet_raw_posterior(const raw_table signal, float min_prob,
                                      float tempW, float tempb, bool return_log);
scrappie_matrix nanonet_rgrgr_r94_posterior(const raw_table signal, float min_prob,
                                            float tempW, float tempb, bool return_log);
scrappie_matrix nanonet_rgrgr_r941_posterior(const raw_table signal, float min_prob,
                                             float tempW, float tempb, bool return_log);
scrappie_matrix nanonet_rgrgr_r10_posterior(const raw_table signal, float min_prob,
                                            float tempW, float tempb, bool return_log);
scrappie_matrix nanonet_rnnrf_r94_transitions(const raw_table signal, float min_prob,
                                              float tempW, float tempb, bool return_log);

/* -- decoders (GPU): src/decode.c:123-365, :836-893.  seq / path: nblock + 1 ints,
 *    caller allocated.  Return the Viterbi score, NAN on failure. -- */
float decode_transducer(const_scrappie_matrix logpost, float stay_pen, float skip_pen,
                        float local_pen, int *seq, bool allow_slip);
float decode_crf(const_scrappie_matrix trans, int *path);
/* python/pyscrap.h:44-61, src/decode.h:40-49 (src/decode.c:1420-1964, :1638): alignment of a log posterior to a
   known k-mer sequence -- Viterbi (optionally with the path: nblock ints, -1 = start / end state) or forward score,
   full or banded (poslow inclusive, poshigh exclusive, one pair per block).  NAN on failure.  The DP runs on the GPU.
   encode_bases_to_integers (src/scrappie_seq_helpers.c:53-75): calloc'd array of n - state_len + 1 k-mer states. */
float map_to_sequence_viterbi(const_scrappie_matrix logpost, float stay_pen, float skip_pen, float local_pen,
                              int const *seq, size_t seqlen, int *path);
float map_to_sequence_forward(const_scrappie_matrix logpost, float stay_pen, float skip_pen, float local_pen,
                              int const *seq, size_t seqlen);
float map_to_sequence_viterbi_banded(const_scrappie_matrix logpost, float stay_pen, float skip_pen, float local_pen,
                                     int const *seq, size_t seqlen, size_t const *poslow, size_t const *poshigh);
float map_to_sequence_forward_banded(const_scrappie_matrix logpost, float stay_pen, float skip_pen, float local_pen,
                                     int const *seq, size_t seqlen, size_t const *poslow, size_t const *poshigh);
bool are_bounds_sane(size_t const *low, size_t const *high, size_t nblock, size_t seqlen);
int *encode_bases_to_integers(char const *seq, size_t n, size_t state_len);

/* python/pyscrap.h:22, src/decode.h:30 (src/decode.c:928-1012): per-block state probabilities (ACGT-) of a CRF;
   new 5 x (nblock + 1) matrix owned by the caller, NULL on failure */
scrappie_matrix posterior_crf(const_scrappie_matrix trans);

/* -- integer post-processing (host): src/decode.c:449-509, :895-918,
 *    src/homopolymer.c:175-235.  Returned strings are calloc'd; caller frees. -- */
char *overlapper(const int *seq, size_t n, int nkmer, int *pos);
char *crfpath_to_basecall(int const *path, size_t npos, int *pos);
int homopolymer_path(const_scrappie_matrix post, int *viterbipath,
                     enum homopolymer_calculation pathCalculationFlag);
enum homopolymer_calculation get_homopolymer_calculation(const char *calcstr);

/* ===================================================================== */
/* Part 2 -- batch interface (new)                                        */
/* ===================================================================== */

typedef struct sb2_engine sb2_engine;   /* one CUDA device + resident model weights */
typedef struct sb2_batch sb2_batch;     /* device workspace for one batch of reads   */

/* Decode / network parameters; sb2_default_params() = `scrappie raw` defaults
 * (src/scrappie_raw.c:98-121). */
typedef struct {
    float min_prob, tempW, tempb;
    float stay_pen, skip_pen, local_pen;
    int allow_slip;
    int homopolymer;            /* enum homopolymer_calculation */
} sb2_params;
sb2_params sb2_default_params(void);

/* Engine for `device`.  Weights are read from `weights_dir` (NULL: $SCRAPPIE_B200_WEIGHTS,
 * else the `weights/` directory next to the shared library).  NULL on failure. */
sb2_engine *sb2_engine_create(int device, const char *weights_dir);
void sb2_engine_destroy(sb2_engine *eng);
/* Install a model from a weight blob already in host memory (multi-GPU runs broadcast
 * the blob from rank 0 and never touch the file system on the other ranks). */
int sb2_engine_load_blob(sb2_engine *eng, enum raw_model_type model, const void *blob, size_t nbytes);
/* Last error message of the calling thread ("" if none). */
const char *sb2_last_error(void);
/* Number of kernel launches this engine has issued so far. */
uint64_t sb2_engine_launch_count(const sb2_engine *eng);

/* A batch = nread reads of the given lengths (samples fed to the network, i.e.
 * end - start).  Device buffers are sized here; reuse a batch for equal-or-smaller work
 * by calling sb2_batch_reset. */
sb2_batch *sb2_batch_create(sb2_engine *eng, enum raw_model_type model, const size_t *nsample, size_t nread);
void sb2_batch_destroy(sb2_batch *b);
size_t sb2_batch_nblock(const sb2_batch *b, size_t read);       /* columns of read's posterior */
size_t sb2_batch_total_blocks(const sb2_batch *b);
size_t sb2_batch_nstate(const sb2_batch *b);
/* host -> device copy of the (already trimmed + normalised) signals: one pointer per read,
 * or one buffer in the batch's padded layout (read r starts at sb2_batch_sample_offset(b, r),
 * total sb2_batch_total_samples_padded(b) floats; asynchronous when the buffer is pinned). */
int sb2_batch_upload(sb2_batch *b, const float *const *signals);
int sb2_batch_upload_concat(sb2_batch *b, const float *concat, int pinned_async);
size_t sb2_batch_total_samples_padded(const sb2_batch *b);
size_t sb2_batch_sample_offset(const sb2_batch *b, size_t read);
void *sb2_host_alloc_pinned(size_t nbytes);
void sb2_host_free_pinned(void *p);
/* keep a copy of every layer's output during forward (parity tests only) */
int sb2_batch_keep_layers(sb2_batch *b, int keep);
/* network forward: leaves the posterior / transition matrices in HBM */
int sb2_batch_forward(sb2_batch *b, const sb2_params *p, bool return_log);
/* Viterbi decode of the resident posterior: paths + scores stay in HBM */
int sb2_batch_decode(sb2_batch *b, const sb2_params *p);
/* forward (log posterior) + decode in one call; replays a captured CUDA graph from the third call on */
int sb2_batch_run(sb2_batch *b, const sb2_params *p);
int sb2_batch_sync(sb2_batch *b);
/* posterior_crf for every read of an rnnrf_r94 batch, on the device, after sb2_batch_forward / sb2_batch_run;
   sb2_batch_download_base_probs copies one read's (nblock + 1) x 8 floats (5 used per column) */
int sb2_batch_posterior_crf(sb2_batch *b);
int sb2_batch_download_base_probs(sb2_batch *b, size_t read, float *dst);
/* device -> host */
int sb2_batch_download_posterior(sb2_batch *b, size_t read, float *dst, size_t dst_stride);
int sb2_batch_download_paths(sb2_batch *b, int *paths_concat /* sum(nblock+1) */, float *scores /* nread */);
/* Per-layer activations of one read, for layer-wise parity tests: layer 0 = conv+act,
 * 1..5 = GRU outputs; dst is [nblock][H]. */
int sb2_batch_download_layer(sb2_batch *b, int layer, size_t read, float *dst);
/* Time `nrep` repetitions of forward+decode with CUDA events on the batch's own stream;
 * optionally a buffer larger than L2 is overwritten between repetitions.  ms_out[nrep]. */
int sb2_batch_time(sb2_batch *b, const sb2_params *p, int nrep, int flush_l2, float *ms_out,
                   float *ms_forward_out, float *ms_decode_out);
/* Per-kernel-stage times of the last sb2_batch_time repetition, ms. */
int sb2_batch_stage_ms(const sb2_batch *b, float *stage_ms, int nstage_max);
/* stage boundaries (n <= 15) of the last sb2_multi_time run, ms after the first batch's first launch */
int sb2_batch_stage_offsets(const sb2_batch *b, float *at, int n_max);

/* One basecall, the equivalent of calculate_post (src/scrappie_raw.c:265-315) for many
 * reads: upload, forward, decode, homopolymer fix-up, overlapper.  signals are trimmed +
 * normalised.  Outputs (any may be NULL): bases[i] calloc'd string (caller frees),
 * scores[i], nblock[i]. Returns number of reads called successfully. */
typedef struct {
    char *bases;        /* calloc'd, caller frees; NULL on failure */
    float score;
    size_t nblock;
    size_t nbase;
} sb2_call;
/* Signal preparation of `scrappie raw` (src/scrappie_raw.c:98-121, :270-275): trim_and_segment_raw +
 * medmad_normalise_array, on the device, bit-identical to the host functions of part 1. */
typedef struct {
    size_t trim_start, trim_end, varseg_chunk;
    float varseg_thresh;
} sb2_trim;
sb2_trim sb2_default_trim(void);                     /* 200, 10, 100, 0.0 */
/* start / end per read (both 0 when the read trims to nothing -- the case in which the reference frees it);
 * normalised (may be NULL): normalised[r] receives end[r] - start[r] floats */
int sb2_prepare_reads(sb2_engine *eng, const float *const *raws, const size_t *nsample, size_t nread,
                      const sb2_trim *t, size_t *start, size_t *end, float *const *normalised);
/* calculate_post (src/scrappie_raw.c:265-315) for a batch of untrimmed pA signals: everything from the trimmer to
 * the base strings runs on the device.  start / end (may be NULL) as above; returns the number of reads called. */
int sb2_basecall_raw_batch(sb2_engine *eng, enum raw_model_type model, const float *const *raws,
                           const size_t *nsample, size_t nread, const sb2_trim *t, const sb2_params *p,
                           sb2_call *out, size_t *start, size_t *end);

/* nanonet_posterior for several event tables in one pass (LSTM scans run 8 reads per CTA); out[i] is a new matrix
 * owned by the caller or NULL.  Returns the number of posteriors made, -1 on failure. */
int sb2_events_posterior_batch(sb2_engine *eng, const event_table *tables, size_t ntable, float min_prob,
                               float tempW, float tempb, bool return_log, scrappie_matrix *out);

int sb2_basecall_batch(sb2_engine *eng, enum raw_model_type model, const float *const *signals,
                       const size_t *nsample, size_t nread, const sb2_params *p, sb2_call *out);
void sb2_calls_free(sb2_call *calls, size_t n);     /* free() every calls[i].bases */
/* sb2_basecall_batch / sb2_basecall_raw_batch keep their device workspaces (buffers, pinned staging, CUDA graphs) in a
 * per-engine pool between calls, so a steady stream of calls allocates nothing; concurrent callers each get their own.
 * sb2_engine_trim_pool releases the idle ones (returns how many); sb2_engine_destroy releases all. */
int sb2_engine_trim_pool(sb2_engine *eng);
/* how often a workspace (device buffers, pinned staging, base-string area) was (re)allocated so far: constant once the
 * pool is warm */
uint64_t sb2_engine_realloc_count(const sb2_engine *eng);
/* Fault injection for tests of the error paths (the reference tests its own with a failing allocator,
 * src/scrappie_stdlib.h:10-37): the workspace allocation number `nth` from now on (0 = the next one; device and pinned
 * host allocations of batch workspaces both count) fails with "injected allocation failure"; a negative value switches
 * the hook off.  Every entry point must then return its error value (NULL / NAN / -1) with sb2_last_error set, release
 * what the failed call had allocated and leave the engine -- and a caller-owned batch -- usable.  Process-wide. */
void sb2_debug_fail_alloc(long nth);
/* Same on an existing batch workspace (no device allocation per call).  concat: signals in
 * the batch's padded layout (pinned != 0 if it came from sb2_host_alloc_pinned), or NULL
 * when the signals are already resident. */
int sb2_batch_basecall(sb2_batch *b, const float *concat, int pinned, const sb2_params *p, sb2_call *out);
/* Time forward+decode of several batches running concurrently on their own streams (how a
 * job larger than one batch executes); CUDA events on the launching stream. */
int sb2_multi_time(sb2_batch **batches, int nbatch, const sb2_params *p, int nrep, int flush_l2, float *ms_out);
/* streaming throughput: batch k runs nrep_per_batch[k] steps back to back on its own stream, no synchronisation
   between steps or batches; *ms_total = device time of all the runs */
int sb2_multi_stream_time(sb2_batch **batches, int nbatch, const sb2_params *p, const int *nrep_per_batch, float *ms_total);

/* Diagnostic hook (SCRAPPIE_B200_TRACE=1 at engine creation): clock64() stamps recorded by CTA 0 of
 * the second GRU layer's scan at its hand-over points, steps 100..103, 16 slots per step. */
int sb2_engine_read_trace(sb2_engine *eng, long long *out, int n);

/* Test hook: flat dump of the host-side convolution tail plan, which reproduces the
 * right-edge behaviour of the reference's strided convolution (src/layers.c:218-241).
 * out = {first_col, ncol, then 24 x {nseg, (x0, tap0, ntap) x 3}}. */
int sb2_conv_plan_debug(size_t nsample, size_t winlen, size_t stride, int *out, int nout);

#ifdef __cplusplus
}
#endif
#endif /* SCRAPPIE_B200_H */
