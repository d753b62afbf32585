/* Internal declarations shared by the host C files and the CUDA translation units. */
#ifndef SB2_INTERNAL_H
#define SB2_INTERNAL_H

#include <stddef.h>
#include <stdint.h>

#include "scrappie_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SB2_NLAYER 5
#define SB2_NMODEL ((int)SCRAPPIE_MODEL_INVALID)

/* ---- error reporting (thread local) ---- */
void sb2_set_error(const char *fmt, ...);

/* ---- weight blob (layout: tools/extract_weights.py) ---- */
typedef struct {
    const float *data;      /* host pointer into the blob: nc columns of `stride` floats */
    uint32_t nr, nc, stride;
} sb2_tensor;

typedef struct {
    void *blob;             /* owned copy of the blob */
    size_t nbytes;
    uint32_t conv_stride, conv_act, head, residual;
    uint32_t arch;              /* 0: conv + 5 alternating GRU layers + head; 1: raw_r94 (bidirectional pairs) */
    uint32_t winlen, H, nstate, ostride;
    uint32_t nfilter;           /* convolution filters (= H for arch 0) */
    uint32_t ffw;               /* arch 1: width of the feedforward2_tanh layers that merge a GRU pair */
    sb2_tensor conv_W, conv_b;
    /* arch 0: layers 1..5;  arch 1: entries 0..3 = F1, B1, F2, B2 (src/networks.c:196-247) */
    sb2_tensor iW[SB2_NLAYER], b[SB2_NLAYER], sW[SB2_NLAYER], sW2[SB2_NLAYER];
    sb2_tensor comb_Wf[2], comb_Wb[2], comb_b[2];       /* arch 1: FF1 / FF2 */
    sb2_tensor FF_W, FF_b;
} sb2_host_model;

int sb2_host_model_parse(const void *blob, size_t nbytes, sb2_host_model *out);  /* copies blob */
void sb2_host_model_free(sb2_host_model *m);
int sb2_read_file(const char *path, void **data, size_t *nbytes);
const char *sb2_model_file_stem(enum raw_model_type model);
/* directory holding <model>.bin; buf receives the path. */
int sb2_default_weights_dir(char *buf, size_t buflen);

/* ---- convolution tail plan --------------------------------------------------
 * The reference's strided convolution (src/layers.c:159-246) equals a zero padded
 * "same" convolution except in its last few output columns.  Columns >= first_col are
 * described explicitly: out[col] = bias + sum over segments of
 * dot(taps[tap0 .. tap0+ntap), x[x0 .. x0+ntap)). */
#define SB2_CONV_TAIL_COLS 24
#define SB2_CONV_TAIL_SEGS 3
typedef struct {
    int32_t first_col;
    int32_t ncol;                                   /* total output columns of the read */
    int32_t nseg[SB2_CONV_TAIL_COLS];
    int32_t seg[SB2_CONV_TAIL_COLS][SB2_CONV_TAIL_SEGS][3];   /* x0, tap0, ntap */
} sb2_conv_tail;
int sb2_conv_plan(size_t nsample, size_t winlen, size_t stride, sb2_conv_tail *plan);

/* ---- homopolymer runs (host_decode.c) ---- */
typedef struct { int start, length, state; } sb2_hp_run;
int sb2_find_homopolymer_runs(const int *path, int pathlen, int klen, sb2_hp_run **runs_out);
void sb2_apply_homopolymer_run(int *path, const sb2_hp_run *run, const float *logp_stay,
                               const float *logp_rep);

#ifdef __cplusplus
}
#endif
#endif
