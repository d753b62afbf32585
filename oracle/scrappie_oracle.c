/* TEST INFRASTRUCTURE ONLY -- see scrappie_oracle.h.
 *
 * Scalar C restatement of the reference's `scrappie raw` hot path.  Compiled with
 * -ffp-contract=off so every multiply/add rounds separately, as the reference's
 * -std=c99 build does.  Dense products are plain left-to-right dot products; the
 * reference delegates those to an (unpinned) cblas, so summation order -- and only
 * that -- differs from oracle/_ref.
 */
#include "scrappie_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* scalar maths                                                              */
/* ------------------------------------------------------------------------- */

static inline float bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* exp_ps, src/sse_mathfun.h:211-290 (cephes polynomial, input clamped to +-88.376) */
float sb2o_expf(float x) {
    const float hi = 88.3762626647949f, lo = -88.3762626647949f;
    x = (x < hi) ? x : hi;
    x = (x > lo) ? x : lo;
    float fx = x * (float)1.44269504088896341;
    fx = fx + 0.5f;
    float fl = (float)(int32_t)fx;      /* truncate, then fix up to floor */
    if (fl > fx) fl = fl - 1.0f;
    fx = fl;
    float t = fx * (float)0.693359375;
    float z = fx * (float)-2.12194440e-4;
    x = x - t;
    x = x - z;
    z = x * x;
    float y = (float)1.9875691500E-4;
    y = y * x; y = y + (float)1.3981999507E-3;
    y = y * x; y = y + (float)8.3334519073E-3;
    y = y * x; y = y + (float)4.1665795894E-2;
    y = y * x; y = y + (float)1.6666665459E-1;
    y = y * x; y = y + (float)5.0000001201E-1;
    y = y * z;
    y = y + x;
    y = y + 1.0f;
    const float pow2n = bits2f((uint32_t)(((int32_t)fx + 127) << 23));
    return y * pow2n;
}

/* log_ps, src/sse_mathfun.h:123-205 */
float sb2o_logf(float x) {
    const int invalid = (x <= 0.0f);
    const float min_norm = bits2f(0x00800000u);
    x = (x > min_norm) ? x : min_norm;
    int32_t e_i = (int32_t)(f2bits(x) >> 23);
    x = bits2f((f2bits(x) & ~0x7f800000u) | f2bits(0.5f));
    e_i -= 127;
    float e = (float)e_i;
    e = e + 1.0f;
    const int small = (x < (float)0.707106781186547524);
    const float tmp0 = small ? x : 0.0f;
    x = x - 1.0f;
    e = e - (small ? 1.0f : 0.0f);
    x = x + tmp0;
    const float z = x * x;
    float y = (float)7.0376836292E-2;
    y = y * x; y = y + (float)-1.1514610310E-1;
    y = y * x; y = y + (float)1.1676998740E-1;
    y = y * x; y = y + (float)-1.2420140846E-1;
    y = y * x; y = y + (float)1.4249322787E-1;
    y = y * x; y = y + (float)-1.6668057665E-1;
    y = y * x; y = y + (float)2.0000714765E-1;
    y = y * x; y = y + (float)-2.4999993993E-1;
    y = y * x; y = y + (float)3.3333331174E-1;
    y = y * x;
    y = y * z;
    float tmp = e * (float)-2.12194440e-4;
    y = y + tmp;
    tmp = z * 0.5f;
    y = y - tmp;
    tmp = e * (float)0.693359375;
    x = x + y;
    x = x + tmp;
    return invalid ? NAN : x;
}

/* logisticfv / tanhfv / elufv, src/util.h:180-198 */
float sb2o_logisticf(float x) { return 1.0f / (1.0f + sb2o_expf(-x)); }
float sb2o_tanhf(float x) { const float y = sb2o_logisticf(x + x); return (y + y) - 1.0f; }
float sb2o_eluf(float x) { return (x >= 0.0f) ? x : (sb2o_expf(x) - 1.0f); }
/* src/util.h:162-164 (libm expf/log1pf) */
float sb2o_logsumexpf(float x, float y) { return fmaxf(x, y) + log1pf(expf(-fabsf(x - y))); }

/* ------------------------------------------------------------------------- */
/* weight blob                                                               */
/* ------------------------------------------------------------------------- */

typedef struct { char name[24]; uint32_t nr, nc, stride, offset; } blob_entry;

static int find_tensor(const blob_entry *tab, uint32_t n, const float *data, const char *name,
                       sb2o_tensor *out) {
    for (uint32_t i = 0; i < n; i++) {
        if (0 == strncmp(tab[i].name, name, sizeof(tab[i].name))) {
            out->data = data + tab[i].offset;
            out->nr = tab[i].nr; out->nc = tab[i].nc; out->stride = tab[i].stride;
            return 0;
        }
    }
    return -1;
}

int sb2o_model_from_blob(const void *blob, size_t nbytes, sb2o_model *m) {
    if (NULL == blob || NULL == m || nbytes < 40) return -1;
    const unsigned char *p = blob;
    if (0 != memcmp(p, "SB2WTS01", 8)) return -1;
    uint32_t hdr[8];
    memcpy(hdr, p + 8, sizeof(hdr));
    const uint32_t nt = hdr[0];
    m->conv_stride = hdr[1]; m->conv_act = hdr[2]; m->head = hdr[3]; m->residual = hdr[4]; m->arch = hdr[5];
    const blob_entry *tab = (const blob_entry *)(p + 40);
    const float *data = (const float *)(p + 40 + (size_t)nt * sizeof(blob_entry));
    int rc = 0;
    if (m->arch != 2) {
        rc |= find_tensor(tab, nt, data, "conv_W", &m->conv_W);
        rc |= find_tensor(tab, nt, data, "conv_b", &m->conv_b);
    }
    if (m->arch >= 1) {
        const char *cn[2][3] = {{"comb1_Wf", "comb1_Wb", "comb1_b"}, {"comb2_Wf", "comb2_Wb", "comb2_b"}};
        for (int i = 0; i < 2; i++) {
            rc |= find_tensor(tab, nt, data, cn[i][0], &m->comb_Wf[i]);
            rc |= find_tensor(tab, nt, data, cn[i][1], &m->comb_Wb[i]);
            rc |= find_tensor(tab, nt, data, cn[i][2], &m->comb_b[i]);
        }
    }
    for (int l = 0; l < (m->arch >= 1 ? 4 : 5); l++) {
        char nm[24];
        const char *parts[4] = {"iW", "b", "sW", "sW2"};
        sb2o_tensor *dst[4] = {&m->iW[l], &m->b[l], &m->sW[l], &m->sW2[l]};
        for (int k = 0; k < 4; k++) {
            strcpy(nm, "gru0_");
            nm[3] = (char)('1' + l);
            strcat(nm, parts[k]);
            rc |= find_tensor(tab, nt, data, nm, dst[k]);
        }
    }
    rc |= find_tensor(tab, nt, data, "FF_W", &m->FF_W);
    rc |= find_tensor(tab, nt, data, "FF_b", &m->FF_b);
    return rc;
}

size_t sb2o_nstate(const sb2o_model *m) { return m->FF_W.nc; }

/* ------------------------------------------------------------------------- */
/* layers                                                                    */
/* ------------------------------------------------------------------------- */

size_t sb2o_conv_ncol(size_t nsample, size_t stride) { return (nsample + stride - 1) / stride; }

/* One BLAS call of the reference's convolution contributes, to output column
 * `col`, the dot product of taps [tap0, tap0+ntap) with samples [x0, x0+ntap). */
static void conv_add(const float *x, size_t n, const sb2o_tensor *W, size_t col, size_t ncol,
                     size_t tap0, size_t x0, size_t ntap, float *out) {
    if (col >= ncol) return;                    /* the reference would write out of bounds */
    const size_t nf = W->nc;
    for (size_t f = 0; f < nf; f++) {
        const float *w = W->data + f * W->stride;   /* taps live at every 4th float */
        float acc = 0.0f;
        for (size_t k = 0; k < ntap; k++) {
            const float xv = (x0 + k < n) ? x[x0 + k] : 0.0f;
            acc += w[4 * (tap0 + k)] * xv;
        }
        out[col * nf + f] += acc;
    }
}

/* convolution(), src/layers.c:159-246, for a single input feature (raw signal).
 * The three call groups of the reference are enumerated with its own index
 * arithmetic (:190-196 left edge, :209-224 strided body, :227-241 right edge), so
 * the stride>1 tail behaviour (body drops the last complete window when
 * (n - shift - w) is a multiple of nstepX; right-edge columns land where
 * offsetC_R says) is reproduced rather than "fixed". */
void sb2o_convolution(const float *x, size_t n, const sb2o_tensor *W, const sb2o_tensor *b,
                      size_t stride, float *out) {
    const size_t winlen = W->stride / 4;            /* W->nrq / X->nrq with X->nrq == 1 */
    const size_t nf = W->nc;
    const size_t padL = (winlen - 1) / 2, padR = winlen / 2;
    const size_t ncol = sb2o_conv_ncol(n, stride);
    for (size_t c = 0; c < ncol; c++) memcpy(out + c * nf, b->data, nf * sizeof(float));

    for (size_t w = 0; w < padL; w += stride)       /* :190-196 */
        conv_add(x, n, W, w / stride, ncol, padL - w, 0, winlen - (padL - w), out);

    const size_t ncolL = (padL + stride - 1) / stride;      /* :199 */
    const size_t shiftL = ncolL * stride - padL;            /* :203 */
    const size_t nstepC = (winlen + stride - 1) / stride;   /* :206 */
    const size_t nstepX = stride * nstepC;                  /* :207 */
    for (size_t w = 0; w < winlen; w += stride) {           /* :209-224 */
        const size_t nproc = ((long)n - (long)shiftL - (long)w > 0) ? (n - shiftL - w) / nstepX : 0;
        const size_t col0 = w / stride + ncolL;
        for (size_t j = 0; j < nproc; j++)
            conv_add(x, n, W, col0 + j * nstepC, ncol, 0, shiftL + w + j * nstepX, winlen, out);
    }

    const size_t maxcol = (n - shiftL) / nstepX;            /* :227 */
    const size_t rem = (n - shiftL) % nstepX;               /* :228 */
    const long colR = (long)ncolL + (long)nstepC * ((long)maxcol - 1) + (long)(rem / stride) + 1;
    const size_t xR = n - winlen + 1;                       /* :232 */
    const long startR = (long)stride - (long)((padL + n - winlen) % stride) - 1;   /* :234 */
    for (long w = startR; w < (long)padR; w += (long)stride) {   /* :235-241 */
        const long col = colR + w / (long)stride;
        if (col < 0) continue;
        conv_add(x, n, W, (size_t)col, ncol, 0, xR + (size_t)w, winlen - 1 - (size_t)w, out);
    }
}

/* affine_map, src/scrappie_matrix.c:323-351: out[c][k] = b[k] + sum_i W[k][i] X[c][i] */
void sb2o_affine(const float *X, size_t ncol, const sb2o_tensor *W, const sb2o_tensor *b, float *out) {
    const size_t nin = W->nr, nout = W->nc;
    for (size_t c = 0; c < ncol; c++) {
        const float *xc = X + c * nin;
        for (size_t k = 0; k < nout; k++) {
            const float *w = W->data + k * W->stride;
            float acc = 0.0f;
            for (size_t i = 0; i < nin; i++) acc += w[i] * xc[i];
            out[c * nout + k] = b->data[k] + acc;
        }
    }
}

/* gru_forward / gru_backward / gru_step, src/layers.c:373-527.
 * Gate order in Xin / sW columns: [0,H) update z (keeps the OLD state), [H,2H) reset r,
 * [2H,3H) candidate.  Reset is applied to the state BEFORE the second product. */
void sb2o_gru(const float *Xin, size_t ncol, const sb2o_tensor *sW, const sb2o_tensor *sW2,
              int backward, float *out) {
    const size_t H = sW2->nc;
    float *h = calloc(H, sizeof(float));
    float *g = malloc(3 * H * sizeof(float));
    float *rh = malloc(H * sizeof(float));
    for (size_t s = 0; s < ncol; s++) {
        const size_t t = backward ? (ncol - 1 - s) : s;
        const float *x = Xin + t * 3 * H;
        for (size_t k = 0; k < 2 * H; k++) {
            const float *w = sW->data + k * sW->stride;
            float acc = 0.0f;
            for (size_t i = 0; i < H; i++) acc += w[i] * h[i];
            g[k] = sb2o_logisticf(x[k] + acc);
        }
        for (size_t i = 0; i < H; i++) rh[i] = g[H + i] * h[i];
        for (size_t k = 0; k < H; k++) {
            const float *w = sW2->data + k * sW2->stride;
            float acc = 0.0f;
            for (size_t i = 0; i < H; i++) acc += w[i] * rh[i];
            g[2 * H + k] = sb2o_tanhf(x[2 * H + k] + acc);
        }
        float *o = out + t * H;
        for (size_t i = 0; i < H; i++) {
            const float z = g[i];
            o[i] = z * h[i] + (1.0f - z) * g[2 * H + i];
        }
        memcpy(h, o, H * sizeof(float));
    }
    free(rh); free(g); free(h);
}

/* softmax_with_temperature + robustlog, src/layers.c:340-357, :79-94,
 * row_normalise_inplace src/scrappie_matrix.c:385-407 (4 lane-wise partial sums,
 * padding lanes of the last quad removed, then a pairwise horizontal add). */
static void head_softmax(const float *X, size_t ncol, const sb2o_tensor *W, const sb2o_tensor *b,
                         float min_prob, float tempW, float tempb, int return_log, float *out) {
    const size_t nin = W->nr, ns = W->nc, ostride = 4 * ((ns + 3) / 4);
    float *xs = malloc(nin * sizeof(float));
    for (size_t c = 0; c < ncol; c++) {
        float *o = out + c * ostride;
        for (size_t i = 0; i < nin; i++) xs[i] = (X[c * nin + i] - 0.0f) / (tempW / tempb);
        for (size_t k = 0; k < ostride; k++) {
            float v = 0.0f;                         /* padding rows: bias padding is zero */
            if (k < ns) {
                const float *w = W->data + k * W->stride;
                float acc = 0.0f;
                for (size_t i = 0; i < nin; i++) acc += w[i] * xs[i];
                v = ((b->data[k] + acc) - 0.0f) / tempb;
            }
            o[k] = sb2o_expf(v);
        }
        float lane[4] = {0, 0, 0, 0};
        for (size_t k = 0; k < ostride; k++) {
            if (k < 4) lane[k] = o[k]; else lane[k & 3] += o[k];
        }
        for (size_t k = ns; k < ostride; k++) lane[k & 3] -= o[k];
        const float tsum = (lane[0] + lane[1]) + (lane[2] + lane[3]);
        const float recip = 1.0f / tsum;
        for (size_t k = 0; k < ostride; k++) o[k] *= recip;
        if (return_log)
            for (size_t k = 0; k < ostride; k++)
                o[k] = sb2o_logf(min_prob + (1.0f - min_prob) * o[k]);
    }
    free(xs);
}

/* globalnorm + crf_partition_function, src/layers.c:835-889 */
static void head_globalnorm(const float *X, size_t ncol, const sb2o_tensor *W, const sb2o_tensor *b,
                            float *out) {
    const size_t ns2 = W->nc, ostride = 4 * ((ns2 + 3) / 4);
    const size_t ns = (size_t)roundf(sqrtf((float)ns2));
    float *tmp = malloc(ncol * ns2 * sizeof(float));
    sb2o_affine(X, ncol, W, b, tmp);
    float prev[16] = {0}, curr[16] = {0};
    for (size_t c = 0; c < ncol; c++) {
        memcpy(prev, curr, sizeof(prev));
        for (size_t to = 0; to < ns; to++) {
            const float *row = tmp + c * ns2 + to * ns;
            float v = row[0] + prev[0];
            for (size_t from = 1; from < ns; from++) v = sb2o_logsumexpf(v, row[from] + prev[from]);
            curr[to] = v;
        }
    }
    float logZ = curr[0];
    for (size_t s = 1; s < ns; s++) logZ = sb2o_logsumexpf(logZ, curr[s]);
    logZ = logZ / (float)ncol;
    for (size_t c = 0; c < ncol; c++) {
        for (size_t k = 0; k < ostride; k++) out[c * ostride + k] = 0.0f;
        for (size_t k = 0; k < ns2; k++) out[c * ostride + k] = tmp[c * ns2 + k] - logZ;
    }
    free(tmp);
}

/* feedforward2_tanh -> affine_map2 (src/layers.c:359-371, src/scrappie_matrix.c:353-383):
 * out[c][k] = tanh((b[k] + sum Wf[k][i] Xf[c][i]) + sum Wb[k][i] Xb[c][i]) */
static void affine2_tanh(const float *Xf, const float *Xb, size_t ncol, const sb2o_tensor *Wf,
                         const sb2o_tensor *Wb, const sb2o_tensor *b, float *out) {
    const size_t nin = Wf->nr, nout = Wf->nc;
    for (size_t c = 0; c < ncol; c++)
        for (size_t k = 0; k < nout; k++) {
            const float *wf = Wf->data + k * Wf->stride, *wb = Wb->data + k * Wb->stride;
            float af = 0.0f, ab = 0.0f;
            for (size_t i = 0; i < nin; i++) af += wf[i] * Xf[c * nin + i];
            for (size_t i = 0; i < nin; i++) ab += wb[i] * Xb[c * nin + i];
            out[c * nout + k] = sb2o_tanhf((b->data[k] + af) + ab);
        }
}

/* nanonet_raw_posterior, src/networks.c:196-247 */
static size_t posterior_raw_r94(const sb2o_model *m, const float *raw, size_t n, float min_prob,
                                float tempW, float tempb, int return_log, float *out) {
    const size_t NF = m->conv_W.nc, H = m->sW2[0].nc, FW = m->comb_b[0].nr;
    const size_t ncol = sb2o_conv_ncol(n, m->conv_stride);
    const size_t wide = (FW > NF) ? FW : NF;
    float *ff = malloc(ncol * wide * sizeof(float));
    float *ff2 = malloc(ncol * wide * sizeof(float));
    float *xin = malloc(ncol * 3 * H * sizeof(float));
    float *gf = malloc(ncol * H * sizeof(float)), *gb = malloc(ncol * H * sizeof(float));
    sb2o_convolution(raw, n, &m->conv_W, &m->conv_b, m->conv_stride, ff);
    for (size_t i = 0; i < ncol * NF; i++) ff[i] = m->conv_act ? sb2o_tanhf(ff[i]) : sb2o_eluf(ff[i]);
    for (int pair = 0; pair < 2; pair++) {
        sb2o_affine(ff, ncol, &m->iW[2 * pair], &m->b[2 * pair], xin);
        sb2o_gru(xin, ncol, &m->sW[2 * pair], &m->sW2[2 * pair], 0, gf);
        sb2o_affine(ff, ncol, &m->iW[2 * pair + 1], &m->b[2 * pair + 1], xin);
        sb2o_gru(xin, ncol, &m->sW[2 * pair + 1], &m->sW2[2 * pair + 1], 1, gb);
        affine2_tanh(gf, gb, ncol, &m->comb_Wf[pair], &m->comb_Wb[pair], &m->comb_b[pair], ff2);
        float *t = ff; ff = ff2; ff2 = t;
    }
    head_softmax(ff, ncol, &m->FF_W, &m->FF_b, min_prob, tempW, tempb, return_log, out);
    free(gb); free(gf); free(xin); free(ff2); free(ff);
    return ncol;
}

/* ---- events model (src/networks.c:146-194) ------------------------------------------------ */

/* nanonet_features_from_events + studentise_features_kahan, src/nnfeatures.c:47-115.
 * ev: n events x (mean, stdv, length); out: n x 4.  The four SSE lanes of the reference are four
 * independent scalar recurrences here; the reciprocal square root is the same RSQRTPS instruction. */
#include <xmmintrin.h>
void sb2o_event_features(const float *ev, size_t n, float *out) {
    for (size_t e = 0; e < n; e++) {
        out[4 * e + 0] = ev[3 * e + 0];
        out[4 * e + 1] = ev[3 * e + 1];
        out[4 * e + 2] = ev[3 * e + 2];
        out[4 * e + 3] = (e + 1 < n) ? (float)fabs(ev[3 * e] - ev[3 * (e + 1)]) : 0.0f;
    }
    float sum[4] = {0}, sumsq[4] = {0}, comp[4] = {0}, compsq[4] = {0};
    for (size_t e = 0; e < n; e++)
        for (int l = 0; l < 4; l++) {
            const float v = out[4 * e + l];
            const float d1 = v - comp[l];
            const float sum_tmp = sum[l] + d1;
            comp[l] = (sum_tmp - sum[l]) - d1;
            sum[l] = sum_tmp;
            const float d2 = v * v - compsq[l];
            const float sumsq_tmp = sumsq[l] + d2;
            compsq[l] = (sumsq_tmp - sumsq[l]) - d2;
            sumsq[l] = sumsq_tmp;
        }
    float var[4];
    for (int l = 0; l < 4; l++) {
        sum[l] /= (float)(int)n;
        sumsq[l] /= (float)(int)n;
        var[l] = sumsq[l] - sum[l] * sum[l];
    }
    float rs[4];
    _mm_storeu_ps(rs, _mm_rsqrt_ps(_mm_loadu_ps(var)));
    for (int l = 0; l < 4; l++) sum[l] *= rs[l];
    for (size_t e = 0; e < n; e++)
        for (int l = 0; l < 4; l++) out[4 * e + l] = rs[l] * out[4 * e + l] - sum[l];
}

/* window(features, 3, 1), src/layers.c:119-146: column c = [f[c-1], f[c], f[c+1]], zero beyond the end.
 * Column 0 stays ALL ZERO in the reference: its loop `for (int w1 = icol - wh + 1; w1 <= icol + wh; ...)`
 * compares the int w1 = -1 with a size_t, i.e. as a huge unsigned value, and never runs.  (The loop also
 * visits a fourth position icol + 2, which lands in the next column's first rows and is overwritten there.) */
static void window3(const float *f, size_t n, float *out) {
    for (size_t c = 0; c < n; c++)
        for (int w = -1; w <= 1; w++)
            for (int l = 0; l < 4; l++) {
                const long src = (long)c + w;
                out[12 * c + 4 * (w + 1) + l] = (0 == c || src >= (long)n) ? 0.0f : f[4 * src + l];
            }
}

/* lstm_forward / lstm_backward / lstm_step, src/layers.c:673-832.  Rows of Xin / columns of sW:
 * [0,H) cell input (tanh), [H,2H) input gate, [2H,3H) forget gate, [3H,4H) output gate;
 * peepholes p: [0,H) input gate, [H,2H) forget gate (both see the OLD cell), [2H,3H) output gate (NEW cell). */
void sb2o_lstm(const float *Xin, size_t ncol, const sb2o_tensor *sW, const sb2o_tensor *p, int backward, float *out) {
    const size_t H = sW->nr;
    float *h = calloc(H, sizeof(float)), *c = calloc(H, sizeof(float));
    float *xF = malloc(4 * H * sizeof(float));
    for (size_t s = 0; s < ncol; s++) {
        const size_t t = backward ? (ncol - 1 - s) : s;
        const float *x = Xin + t * 4 * H;
        for (size_t k = 0; k < 4 * H; k++) {
            const float *w = sW->data + k * sW->stride;
            float acc = 0.0f;
            for (size_t i = 0; i < H; i++) acc += w[i] * h[i];
            xF[k] = x[k] + acc;
        }
        float *o = out + t * H;
        for (size_t i = 0; i < H; i++) {
            const float forget = sb2o_logisticf(xF[2 * H + i] + c[i] * p->data[H + i]) * c[i];
            const float update = sb2o_logisticf(xF[H + i] + c[i] * p->data[i]) * sb2o_tanhf(xF[i]);
            c[i] = forget + update;
            o[i] = sb2o_logisticf(xF[3 * H + i] + c[i] * p->data[2 * H + i]) * sb2o_tanhf(c[i]);
        }
        memcpy(h, o, H * sizeof(float));
    }
    free(xF); free(c); free(h);
}

/* nanonet_posterior, src/networks.c:146-194.  ev: n x (mean, stdv, length); out: n columns of out_stride floats */
size_t sb2o_events_posterior(const sb2o_model *m, const float *ev, size_t n, float min_prob, float tempW,
                             float tempb, int return_log, float *out) {
    if (NULL == m || NULL == ev || 0 == n || NULL == out || m->arch != 2) return 0;
    const size_t H = m->sW[0].nr, FW = m->comb_b[0].nr;
    float *feat = malloc(n * 4 * sizeof(float)), *f3 = malloc(n * 12 * sizeof(float));
    float *ff = malloc(n * FW * sizeof(float)), *ff2 = malloc(n * FW * sizeof(float));
    float *xin = malloc(n * 4 * H * sizeof(float));
    float *lf = malloc(n * H * sizeof(float)), *lb = malloc(n * H * sizeof(float));
    sb2o_event_features(ev, n, feat);
    window3(feat, n, f3);
    const float *in = f3;
    for (int pair = 0; pair < 2; pair++) {
        sb2o_affine(in, n, &m->iW[2 * pair], &m->b[2 * pair], xin);
        sb2o_lstm(xin, n, &m->sW[2 * pair], &m->sW2[2 * pair], 0, lf);
        sb2o_affine(in, n, &m->iW[2 * pair + 1], &m->b[2 * pair + 1], xin);
        sb2o_lstm(xin, n, &m->sW[2 * pair + 1], &m->sW2[2 * pair + 1], 1, lb);
        float *dst = (pair == 0) ? ff : ff2;
        affine2_tanh(lf, lb, n, &m->comb_Wf[pair], &m->comb_Wb[pair], &m->comb_b[pair], dst);
        in = dst;
    }
    head_softmax(in, n, &m->FF_W, &m->FF_b, min_prob, tempW, tempb, return_log, out);
    free(lb); free(lf); free(xin); free(ff2); free(ff); free(f3); free(feat);
    return n;
}

/* nanonet_rgrgr_*_posterior (src/networks.c:250-296) / nanonet_rnnrf_r94_transitions (:567-615) */
size_t sb2o_posterior(const sb2o_model *m, const float *raw, size_t n, float min_prob,
                      float tempW, float tempb, int return_log, float *out, float **layer_out) {
    if (NULL == m || NULL == raw || 0 == n || NULL == out) return 0;
    if (m->arch == 1) return posterior_raw_r94(m, raw, n, min_prob, tempW, tempb, return_log, out);
    const size_t H = m->conv_W.nc;
    const size_t ncol = sb2o_conv_ncol(n, m->conv_stride);
    float *cur = malloc(ncol * H * sizeof(float));
    float *nxt = malloc(ncol * H * sizeof(float));
    float *xin = malloc(ncol * 3 * H * sizeof(float));
    sb2o_convolution(raw, n, &m->conv_W, &m->conv_b, m->conv_stride, cur);
    for (size_t i = 0; i < ncol * H; i++)
        cur[i] = m->conv_act ? sb2o_tanhf(cur[i]) : sb2o_eluf(cur[i]);
    if (layer_out && layer_out[0]) memcpy(layer_out[0], cur, ncol * H * sizeof(float));
    for (int l = 0; l < 5; l++) {
        sb2o_affine(cur, ncol, &m->iW[l], &m->b[l], xin);
        sb2o_gru(xin, ncol, &m->sW[l], &m->sW2[l], (l % 2) == 0, nxt);      /* B,F,B,F,B */
        if (m->residual)                                                   /* src/layers.c:303-319 */
            for (size_t i = 0; i < ncol * H; i++) nxt[i] += cur[i];
        float *t = cur; cur = nxt; nxt = t;
        if (layer_out && layer_out[l + 1]) memcpy(layer_out[l + 1], cur, ncol * H * sizeof(float));
    }
    if (m->head == 0)
        head_softmax(cur, ncol, &m->FF_W, &m->FF_b, min_prob, tempW, tempb, return_log, out);
    else
        head_globalnorm(cur, ncol, &m->FF_W, &m->FF_b, out);
    free(xin); free(nxt); free(cur);
    return ncol;
}

/* ------------------------------------------------------------------------- */
/* decoders                                                                  */
/* ------------------------------------------------------------------------- */

#define BIG 1.e30f

/* Maximum (strict, lowest group wins) over the `ngroup` states that share a
 * suffix: best[p] = max_r prev[r * nsuffix + p].  src/decode.c:186-210, :227-251 */
static void suffix_max(const float *prev, int ngroup, int nsuffix, float *best, int *from) {
    for (int p = 0; p < nsuffix; p++) { best[p] = prev[p]; from[p] = p; }
    for (int r = 1; r < ngroup; r++)
        for (int p = 0; p < nsuffix; p++)
            if (best[p] < prev[r * nsuffix + p]) { best[p] = prev[r * nsuffix + p]; from[p] = r * nsuffix + p; }
}

/* decode_transducer + viterbi_local_backtrace, src/decode.c:123-365, :58-98 */
float sb2o_decode_transducer(const float *logpost, size_t nblock, size_t nstate, size_t stride,
                             float stay_pen, float skip_pen, float local_pen, int *seq,
                             int allow_slip) {
    if (NULL == logpost || NULL == seq) return NAN;
    const int nh = (int)nstate - 1, S = nh, E = nh + 1;
    float *score = malloc((nh + 2) * sizeof(float));
    float *prev = malloc((nh + 2) * sizeof(float));
    float *best = malloc(nh * sizeof(float));
    int *from = malloc(nh * sizeof(int));
    int *tb = malloc(nblock * (size_t)(nh + 2) * sizeof(int));
    for (int i = 0; i < nh; i++) score[i] = -BIG;
    score[S] = 0.0f;
    score[E] = -BIG;

    for (size_t blk = 0; blk < nblock; blk++) {
        const float *lp = logpost + blk * stride;
        int *t = tb + blk * (size_t)(nh + 2);
        { float *x = score; score = prev; prev = x; }
        const float stay = lp[nh] - stay_pen;                       /* :174-182 */
        for (int i = 0; i < nh; i++) { score[i] = prev[i] + stay; t[i] = -1; }

        suffix_max(prev, 4, nh / 4, best, from);                    /* step :186-224 */
        for (int i = 0; i < nh; i++) {
            const float s = lp[i] + best[i / 4];
            if (score[i] < s) { score[i] = s; t[i] = from[i / 4]; }
        }
        suffix_max(prev, 16, nh / 16, best, from);                  /* skip :227-270 */
        for (int i = 0; i < nh; i++) {
            const float s = (lp[i] + best[i / 16]) - skip_pen;
            if (score[i] < s) { score[i] = s; t[i] = from[i / 16]; }
        }
        if (allow_slip) {                                           /* slip :273-323 */
            const float slip_pen = (float)(2.0 * skip_pen);
            suffix_max(prev, 64, nh / 64, best, from);
            for (int i = 0; i < nh; i++) {
                const float s = (lp[i] + best[i / 64]) - slip_pen;
                if (score[i] < s) { score[i] = s; t[i] = from[i / 64]; }
            }
        }
        score[S] = prev[S] + fmaxf(-local_pen, lp[nh] - stay_pen); /* :326-336 */
        t[S] = S;
        for (int i = 0; i < nh; i++) {
            const float s = prev[S] + lp[i];
            if (s > score[i]) { score[i] = s; t[i] = S; }
        }
        score[E] = prev[E] + fmax(-local_pen, lp[nh] - stay_pen);  /* :339-349 (double fmax) */
        t[E] = E;
        for (int i = 0; i < nh; i++) {
            const float s = prev[i] - local_pen;
            if (s > score[E]) { score[E] = s; t[E] = i; }
        }
    }

    /* backtrace :58-98; argmaxf (src/util.c:9-23) keeps the first maximum */
    for (size_t i = 0; i <= nblock; i++) seq[i] = -1;
    int last = 0;
    for (int i = 1; i < nh + 2; i++) if (score[i] > score[last]) last = i;
    const float logscore = score[last];
    for (size_t i = 0; i < nblock; i++) {
        const size_t ri = nblock - i - 1;
        const int st = tb[ri * (size_t)(nh + 2) + last];
        if (st >= 0) { seq[ri + 1] = last; last = st; }
    }
    seq[0] = last;
    for (size_t i = 0; i < nblock; i++) { if (seq[i] == S) seq[i] = -1; else break; }
    for (long i = (long)nblock; i >= 0; i--) { if (seq[i] == E) seq[i] = -1; else break; }

    free(tb); free(from); free(best); free(prev); free(score);
    return logscore;
}

/* decode_crf, src/decode.c:836-893.  trans[to * ns + from] */
float sb2o_decode_crf(const float *trans, size_t nblock, size_t stride, int *path) {
    if (NULL == trans || NULL == path) return NAN;
    enum { NS = 5 };
    float cur[NS] = {0}, prev[NS];
    int *tb = malloc(nblock * NS * sizeof(int));
    for (size_t blk = 0; blk < nblock; blk++) {
        const float *tr = trans + blk * stride;
        memcpy(prev, cur, sizeof(prev));
        for (int to = 0; to < NS; to++) {
            float v = tr[to * NS] + prev[0];
            int arg = 0;
            for (int from = 1; from < NS; from++) {
                const float s = tr[to * NS + from] + prev[from];
                if (s > v) { v = s; arg = from; }
            }
            cur[to] = v;
            tb[blk * NS + to] = arg;
        }
    }
    int last = 0;
    for (int i = 1; i < NS; i++) if (cur[i] > cur[last]) last = i;
    const float score = cur[last];
    path[nblock] = last;
    for (size_t blk = nblock; blk > 0; blk--) path[blk - 1] = tb[(blk - 1) * NS + path[blk]];
    free(tb);
    return score;
}

/* posterior_crf, src/decode.c:928-1012: forward-backward over the 5 x 5 transition energies.
 * post: (nblock + 1) columns of post_stride (>= 5) floats; column b holds the normalised state
 * probabilities after block b - 1.  Note the reference's normaliser starts its fold at 0.0f (not
 * at -inf), i.e. the probabilities of a column sum to S / (1 + S) -- kept as is. */
int sb2o_posterior_crf(const float *trans, size_t nblock, size_t stride, float *post, size_t post_stride) {
    if (NULL == trans || NULL == post) return -1;
    enum { NS = 5 };
    for (int st = 0; st < NS; st++) post[st] = 0.0f;
    for (size_t blk = 0; blk < nblock; blk++) {
        const float *tr = trans + blk * stride;
        const float *prev = post + blk * post_stride;
        float *curr = post + (blk + 1) * post_stride;
        for (int st1 = 0; st1 < NS; st1++) {
            curr[st1] = tr[st1 * NS] + prev[0];
            for (int st2 = 1; st2 < NS; st2++) curr[st1] = sb2o_logsumexpf(curr[st1], tr[st1 * NS + st2] + prev[st2]);
        }
    }
    float bufA[NS], bufB[NS];
    float *prev = bufA, *curr = bufB;
    for (int st = 0; st < NS; st++) curr[st] = 0.0f;
    {
        float *last = post + nblock * post_stride;
        float tot = 0.0f;
        for (int st = 0; st < NS; st++) tot = sb2o_logsumexpf(tot, last[st]);
        for (int st = 0; st < NS; st++) last[st] = expf(last[st] - tot);
    }
    for (size_t blk = nblock; blk > 0; blk--) {
        const float *tr = trans + (blk - 1) * stride;
        float *col = post + (blk - 1) * post_stride;
        float *tmp = curr; curr = prev; prev = tmp;
        for (int st = 0; st < NS; st++) curr[st] = tr[st] + prev[0];
        for (int st1 = 1; st1 < NS; st1++)
            for (int st2 = 0; st2 < NS; st2++) curr[st2] = sb2o_logsumexpf(curr[st2], tr[st1 * NS + st2] + prev[st1]);
        float tot = 0.0f;
        for (int st = 0; st < NS; st++) {
            col[st] += curr[st];
            tot = sb2o_logsumexpf(tot, col[st]);
        }
        for (int st = 0; st < NS; st++) col[st] = expf(col[st] - tot);
    }
    return 0;
}

/* map_to_sequence_viterbi / _forward / _viterbi_banded / _forward_banded, src/decode.c:1420-1964.
 * One restatement with two switches: `forward` replaces max by logsumexpf, `low/high` (may be NULL) are the
 * band limits per block.  The banded variants only rewrite the positions inside the band of each block, so
 * positions outside it keep the value they had two blocks earlier (the two score vectors alternate) -- kept.
 * path (viterbi, unbanded only; may be NULL): nblock ints, -1 for the start / end states. */
float sb2o_map_to_sequence(const float *lp, size_t nblock, size_t nst, size_t stride, float stay_pen, float skip_pen,
                           float local_pen, const int *seq, size_t seqlen, int forward, const size_t *low,
                           const size_t *high, int *path) {
    if (NULL == lp || NULL == seq || seqlen < 3 || 0 == nblock) return NAN;
    const size_t STAY = nst - 1, START = seqlen, END = seqlen + 1, ns = seqlen + 2;
    const int banded = (NULL != low && NULL != high);
    float *c = calloc(ns, sizeof(float)), *p = calloc(ns, sizeof(float));
    int *tb = (!forward && !banded && NULL != path) ? malloc(nblock * ns * sizeof(int)) : NULL;
#define SB2O_COMB(a, b) (forward ? sb2o_logsumexpf((a), (b)) : fmaxf((a), (b)))
    for (size_t i = 0; i < ns; i++) { c[i] = -1e30f; p[i] = -1e30f; }
    if (banded) p[START] = 0.0f; else c[START] = 0.0f;
    for (size_t blk = 0; blk < nblock; blk++) {
        const float *l = lp + blk * stride;
        int *t = tb ? tb + blk * ns : NULL;
        if (!(banded && 0 == blk)) { float *tmp = p; p = c; c = tmp; }
        const float loc = forward ? sb2o_logsumexpf(-local_pen, l[STAY]) : fmaxf(-local_pen, l[STAY]);
        c[START] = p[START] + loc;
        c[END] = p[END] + loc;
        if (t) { t[START] = (int)START; t[END] = (int)END; }
        if (!banded) {
            for (size_t pos = 0; pos < seqlen; pos++) { c[pos] = p[pos] - stay_pen + l[STAY]; if (t) t[pos] = (int)pos; }
            for (size_t pos = 1; pos < seqlen; pos++) {
                const float sc = p[pos - 1] + l[seq[pos]];
                if (forward) c[pos] = sb2o_logsumexpf(c[pos], sc);
                else if (sc > c[pos]) { c[pos] = sc; if (t) t[pos] = (int)pos - 1; }
            }
            for (size_t pos = 2; pos < seqlen; pos++) {
                const float sc = p[pos - 2] - skip_pen + l[seq[pos]];
                if (forward) c[pos] = sb2o_logsumexpf(c[pos], sc);
                else if (sc > c[pos]) { c[pos] = sc; if (t) t[pos] = (int)pos - 2; }
            }
            const float s0 = p[START] + l[seq[0]];
            if (forward) c[0] = sb2o_logsumexpf(c[0], s0);
            else if (s0 > c[0]) { c[0] = s0; if (t) t[0] = (int)START; }
            const float se = p[seqlen - 1] - local_pen;
            if (forward) c[END] = sb2o_logsumexpf(c[END], se);
            else if (se > c[END]) { c[END] = se; if (t) t[END] = (int)seqlen - 1; }
        } else if (0 == blk) {
            c[0] = SB2O_COMB(c[0], p[0] + l[STAY] - stay_pen);
            if (high[0] > 0) c[1] = l[seq[1]];
            if (high[0] > 1) c[2] = l[seq[2]] - skip_pen;
            c[END] = SB2O_COMB(c[END], p[START] - local_pen);
            c[0] = SB2O_COMB(c[0], p[START] + l[seq[0]]);
            c[END] = SB2O_COMB(c[END], p[seqlen - 1] - local_pen);
        } else {
            for (size_t pos = low[blk]; pos < high[blk - 1]; pos++) c[pos] = p[pos] - stay_pen + l[STAY];
            size_t a = low[blk] > low[blk - 1] + 1 ? low[blk] : low[blk - 1] + 1;
            size_t b = high[blk] < high[blk - 1] + 1 ? high[blk] : high[blk - 1] + 1;
            for (size_t pos = a; pos < b; pos++) c[pos] = SB2O_COMB(p[pos - 1] + l[seq[pos]], c[pos]);
            a = low[blk] > low[blk - 1] + 2 ? low[blk] : low[blk - 1] + 2;
            b = high[blk] < high[blk - 1] + 2 ? high[blk] : high[blk - 1] + 2;
            for (size_t pos = a; pos < b; pos++) c[pos] = SB2O_COMB(p[pos - 2] - skip_pen + l[seq[pos]], c[pos]);
            if (0 == low[blk]) c[0] = SB2O_COMB(c[0], p[START] + l[seq[0]]);
            c[END] = SB2O_COMB(c[END], p[seqlen - 1] - local_pen);
        }
    }
    const float score = SB2O_COMB(c[seqlen - 1], c[END]);
#undef SB2O_COMB
    if (tb) {
        path[nblock - 1] = (c[seqlen - 1] > c[END]) ? (int)seqlen - 1 : (int)END;
        for (size_t blk = nblock - 1; blk > 0; blk--) path[blk - 1] = tb[blk * ns + path[blk]];
        for (size_t blk = 0; blk < nblock; blk++)
            if ((int)START == path[blk] || (int)END == path[blk]) path[blk] = -1;
    }
    free(tb); free(p); free(c);
    return score;
}

static const char BASES[4] = {'A', 'C', 'G', 'T'};

/* overlap(), src/decode.c:367-382 */
static int kmer_overlap(int k1, int k2, int nkmer) {
    int mask = nkmer - 1, ol = 0;
    do { mask >>= 2; k1 &= mask; k2 >>= 2; ol++; } while (k1 != k2);
    return ol;
}

/* overlapper(), src/decode.c:449-509 */
char *sb2o_overlapper(const int *seq, size_t n, int nkmer, int *pos) {
    if (NULL == seq) return NULL;
    size_t klen = 0;
    for (size_t x = (size_t)nkmer; x != 0; x >>= 1) klen++;
    klen /= 2;
    size_t st = 0;
    while (st < n && seq[st] < 0) st++;
    if (st == n) return NULL;
    size_t length = klen;
    int kprev = seq[st];
    for (size_t k = st + 1; k < n; k++) {
        if (seq[k] < 0) continue;
        length += (size_t)kmer_overlap(kprev, seq[k], nkmer);
        kprev = seq[k];
    }
    char *bases = calloc(length + 1, 1);
    for (size_t kmer = (size_t)seq[st], k = 1; k <= klen; k++, kmer >>= 2)
        bases[klen - k] = BASES[kmer & 3];
    if (pos) pos[0] = 0;
    size_t last = klen - 1;
    kprev = seq[st];
    for (size_t k = st + 1; k < n; k++) {
        if (seq[k] < 0) { if (pos) pos[k] = pos[k - 1]; continue; }
        const int ol = kmer_overlap(kprev, seq[k], nkmer);
        if (pos) pos[k] = pos[k - 1] + ol;
        kprev = seq[k];
        size_t kmer = (size_t)seq[k];
        for (int i = 0; i < ol; i++, kmer >>= 2) bases[last + ol - i] = BASES[kmer & 3];
        last += ol;
    }
    return bases;
}

/* crfpath_to_basecall(), src/decode.c:895-918 (pos is never written there) */
char *sb2o_crfpath_to_basecall(const int *path, size_t npos) {
    if (NULL == path) return NULL;
    size_t nb = 0;
    for (size_t i = 0; i < npos; i++) nb += (path[i] < 4);
    char *out = calloc(nb + 1, 1);
    for (size_t i = 0, j = 0; i < npos; i++) if (path[i] < 4) out[j++] = BASES[path[i]];
    return out;
}

static int repeat_kmer(int b, int n) { int y = 0; for (int i = 0; i < n; i++) y = y * 4 + b; return y; }

/* homopolymer_path(MEAN) + findRuns, src/homopolymer.c:67-235.
 * Runs are processed in the order findRuns emits them (base-major, then position). */
int sb2o_homopolymer_path(const float *post, size_t nblock, size_t nstate, size_t stride, int *path) {
    const int plen = (int)nblock;
    const int stay = (int)nstate - 1;
    const int klen = (int)(logf((float)nstate) / logf(4.0f));      /* scrappie_seq_helpers.c:137 */
    const int f1 = 1 << (2 * (klen - 1)), f2 = 1 << (2 * (klen - 2));
    const int cap = plen / 2 > 0 ? plen / 2 : 1;
    int *rs = calloc(cap, sizeof(int)), *rl = calloc(cap, sizeof(int)), *rb = calloc(cap, sizeof(int));
    int nrun = 0;
    for (int base = 0; base < 4; base++) {
        const int rk = repeat_kmer(base, klen), rk1 = repeat_kmer(base, klen - 1), rk2 = repeat_kmer(base, klen - 2);
        for (int i = 1; i < plen - 2; i++) {
            const int p = path[i - 1], q = path[i];
            if ((p % f1 == rk1) && (p != rk) && (p != -1) && (q == -1 || q == rk)) {
                int e = i + 1;
                while (e < plen && (path[e] == -1 || path[e] == rk)) e++;
                rs[nrun] = i; rl[nrun] = e - i; rb[nrun] = base; nrun++;
            }
            if ((p % f2 == rk2) && (p % f1 != rk1) && (p != -1) && (q == -1 || q == rk)) {
                int j = i;
                while (j < plen && path[j] == -1) j++;
                if (path[j] == rk && j < plen - 1) {
                    int e = j + 1;
                    while (e < plen && (path[e] == -1 || path[e] == rk)) e++;
                    rs[nrun] = j; rl[nrun] = e - j; rb[nrun] = base; nrun++;
                }
            }
        }
    }
    for (int r = 0; r < nrun; r++) {
        const int state = repeat_kmer(rb[r], klen);
        const int from = rs[r], to = rs[r] + rl[r] - 1;
        int nvit = 0;
        double nmean = 0.0;
        for (int i = from; i <= to; i++) {
            const double ps = expf(post[(size_t)(i - 1) * stride + stay]);
            const double pr = expf(post[(size_t)(i - 1) * stride + state]);
            nmean += pr / (pr + ps);
            if (path[i] == state) nvit++;
        }
        const int newn = (int)(nmean + 0.5);
        if (newn != nvit)
            for (int i = 0; i <= to - from; i++) path[i + from] = (i < newn) ? state : -1;
    }
    free(rs); free(rl); free(rb);
    return 0;
}

void sb2o_free(void *p) { free(p); }

/* ------------------------------------------------------------------------- */
/* signal preparation                                                        */
/* ------------------------------------------------------------------------- */

static int cmp_float(const void *a, const void *b) {
    const float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

/* quantilef, src/util.c:91-133 (sorted copy, linear interpolation, double arithmetic) */
static float quantile(const float *x, size_t n, float p) {
    float *s = malloc(n * sizeof(float));
    memcpy(s, x, n * sizeof(float));
    qsort(s, n, sizeof(float), cmp_float);
    const size_t idx = (size_t)(p * (n - 1));
    const float rem = p * (n - 1) - idx;
    float q;
    if (idx < n - 1) q = (float)((1.0 - rem) * s[idx] + rem * s[idx + 1]);
    else q = s[idx];
    free(s);
    return q;
}

float sb2o_medianf(const float *x, size_t n) { return quantile(x, n, 0.5f); }

/* madf, src/util.c:160-182 */
float sb2o_madf(const float *x, size_t n, const float *med) {
    const float scale = 1.4826;
    if (1 == n) return 0.0f;
    const float m = med ? *med : sb2o_medianf(x, n);
    float *d = malloc(n * sizeof(float));
    for (size_t i = 0; i < n; i++) d[i] = fabsf(x[i] - m);
    const float mad = sb2o_medianf(d, n);
    free(d);
    return mad * scale;
}

/* medmad_normalise_array, src/util.c:190-204 */
void sb2o_medmad_normalise(float *x, size_t n) {
    if (NULL == x) return;
    if (1 == n) { x[0] = 0.0f; return; }
    const float med = sb2o_medianf(x, n);
    const float mad = sb2o_madf(x, n, &med);
    for (size_t i = 0; i < n; i++) x[i] = (x[i] - med) / mad;
}

/* trim_and_segment_raw + trim_raw_by_mad, src/scrappie_common.c:5-73 (rt.start = 0, rt.end = n on entry) */
int sb2o_trim_and_segment(const float *raw, size_t n, size_t trim_start, size_t trim_end,
                          size_t chunk, float perc, size_t *start_out, size_t *end_out) {
    size_t start = 0, end = n;
    const size_t nchunk = (end - start) / chunk;
    end = nchunk * chunk;
    float *mad = malloc((nchunk ? nchunk : 1) * sizeof(float));
    for (size_t i = 0; i < nchunk; i++) mad[i] = sb2o_madf(raw + start + i * chunk, chunk, NULL);
    const float thresh = nchunk ? quantile(mad, nchunk, perc) : 0.0f;
    for (size_t i = 0; i < nchunk; i++) { if (mad[i] > thresh) break; start += chunk; }
    for (size_t i = nchunk; i > 0; i--) { if (mad[i - 1] > thresh) break; end -= chunk; }
    free(mad);
    start = (n - start) > trim_start ? start + trim_start : n;
    end = (end > trim_end) ? end - trim_end : 0;
    if (start >= end) return -1;
    *start_out = start; *end_out = end;
    return 0;
}
