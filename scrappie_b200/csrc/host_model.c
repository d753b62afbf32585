/* Model registry, weight-blob parsing and the convolution tail plan (host side).
 *
 * Registry functions mirror src/networks.c:17-127 (names, enum, stride, function
 * pointer; errx on an invalid enum, as the reference does).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <err.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sb2_internal.h"

/* ------------------------------------------------------------------ errors */

static __thread char sb2_errbuf[512];

void sb2_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(sb2_errbuf, sizeof(sb2_errbuf), fmt, ap);
    va_end(ap);
    if (NULL != getenv("SCRAPPIE_B200_VERBOSE")) fprintf(stderr, "scrappie_b200: %s\n", sb2_errbuf);
}

const char *sb2_last_error(void) { return sb2_errbuf; }

/* ---------------------------------------------------------------- registry */

static const struct { const char *name; enum raw_model_type type; int stride; } sb2_models[] = {
    {"raw_r94", SCRAPPIE_MODEL_RAW, 5},
    {"rgrgr_r94", SCRAPPIE_MODEL_RGRGR_R9_4, 5},
    {"rgrgr_r941", SCRAPPIE_MODEL_RGRGR_R9_4_1, 5},
    {"rgrgr_r10", SCRAPPIE_MODEL_RGRGR_R10, 5},
    {"rnnrf_r94", SCRAPPIE_MODEL_RNNRF_R9_4, 1},
};
#define SB2_NREG (sizeof(sb2_models) / sizeof(sb2_models[0]))

enum raw_model_type get_raw_model(const char *modelstr) {
    if (NULL == modelstr) return SCRAPPIE_MODEL_INVALID;
    for (size_t i = 0; i < SB2_NREG; i++)
        if (0 == strcmp(modelstr, sb2_models[i].name)) return sb2_models[i].type;
    return SCRAPPIE_MODEL_INVALID;
}

const char *raw_model_string(const enum raw_model_type model) {
    for (size_t i = 0; i < SB2_NREG; i++)
        if (model == sb2_models[i].type) return sb2_models[i].name;
    errx(EXIT_FAILURE, "Invalid scrappie model %s:%d", __FILE__, __LINE__);
    return NULL;
}

int get_raw_model_stride(const enum raw_model_type model) {
    for (size_t i = 0; i < SB2_NREG; i++)
        if (model == sb2_models[i].type) return sb2_models[i].stride;
    errx(EXIT_FAILURE, "Invalid scrappie model %s:%d", __FILE__, __LINE__);
    return -1;
}

/* python/build.py:34-44: -1 for an unknown name instead of exiting */
int get_raw_model_stride_from_string(const char *modelstr) {
    const enum raw_model_type model = get_raw_model(modelstr);
    if (SCRAPPIE_MODEL_INVALID == model) return -1;
    return get_raw_model_stride(model);
}

const char *sb2_model_file_stem(enum raw_model_type model) {
    for (size_t i = 0; i < SB2_NREG; i++)
        if (model == sb2_models[i].type) return sb2_models[i].name;
    return NULL;
}

/* ------------------------------------------------------------ weight blobs */

typedef struct { char name[24]; uint32_t nr, nc, stride, offset; } blob_entry;

static int lookup(const blob_entry *tab, uint32_t n, const float *data, size_t nfloat,
                  const char *name, sb2_tensor *t) {
    for (uint32_t i = 0; i < n; i++) {
        if (0 != strncmp(tab[i].name, name, sizeof(tab[i].name))) continue;
        if ((size_t)tab[i].offset + (size_t)tab[i].stride * tab[i].nc > nfloat) return -1;
        t->data = data + tab[i].offset;
        t->nr = tab[i].nr; t->nc = tab[i].nc; t->stride = tab[i].stride;
        return 0;
    }
    return -1;
}

int sb2_host_model_parse(const void *blob, size_t nbytes, sb2_host_model *m) {
    memset(m, 0, sizeof(*m));
    if (NULL == blob || nbytes < 40 || 0 != memcmp(blob, "SB2WTS01", 8)) {
        sb2_set_error("weight blob: bad magic or truncated");
        return -1;
    }
    m->blob = malloc(nbytes);
    if (NULL == m->blob) return -1;
    memcpy(m->blob, blob, nbytes);
    m->nbytes = nbytes;
    const unsigned char *p = m->blob;
    uint32_t hdr[8];
    memcpy(hdr, p + 8, sizeof(hdr));
    const uint32_t nt = hdr[0];
    const size_t table_bytes = (size_t)nt * sizeof(blob_entry);
    if (40 + table_bytes > nbytes) { sb2_host_model_free(m); return -1; }
    m->conv_stride = hdr[1]; m->conv_act = hdr[2]; m->head = hdr[3]; m->residual = hdr[4]; m->arch = hdr[5];
    const blob_entry *tab = (const blob_entry *)(p + 40);
    const float *data = (const float *)(p + 40 + table_bytes);
    const size_t nfloat = (nbytes - 40 - table_bytes) / sizeof(float);
    if (m->arch > 2) { sb2_set_error("weight blob: unknown architecture %u", m->arch); sb2_host_model_free(m); return -1; }
    const int nlayer = (m->arch == 0) ? SB2_NLAYER : 4;

    int rc = 0;
    if (m->arch != 2) {
        rc |= lookup(tab, nt, data, nfloat, "conv_W", &m->conv_W);
        rc |= lookup(tab, nt, data, nfloat, "conv_b", &m->conv_b);
    }
    for (int l = 0; l < nlayer; l++) {
        char nm[24];
        snprintf(nm, sizeof(nm), "gru%d_iW", l + 1);  rc |= lookup(tab, nt, data, nfloat, nm, &m->iW[l]);
        snprintf(nm, sizeof(nm), "gru%d_b", l + 1);   rc |= lookup(tab, nt, data, nfloat, nm, &m->b[l]);
        snprintf(nm, sizeof(nm), "gru%d_sW", l + 1);  rc |= lookup(tab, nt, data, nfloat, nm, &m->sW[l]);
        snprintf(nm, sizeof(nm), "gru%d_sW2", l + 1); rc |= lookup(tab, nt, data, nfloat, nm, &m->sW2[l]);
    }
    if (m->arch >= 1) {
        for (int i = 0; i < 2; i++) {
            char nm[24];
            snprintf(nm, sizeof(nm), "comb%d_Wf", i + 1); rc |= lookup(tab, nt, data, nfloat, nm, &m->comb_Wf[i]);
            snprintf(nm, sizeof(nm), "comb%d_Wb", i + 1); rc |= lookup(tab, nt, data, nfloat, nm, &m->comb_Wb[i]);
            snprintf(nm, sizeof(nm), "comb%d_b", i + 1);  rc |= lookup(tab, nt, data, nfloat, nm, &m->comb_b[i]);
        }
    }
    rc |= lookup(tab, nt, data, nfloat, "FF_W", &m->FF_W);
    rc |= lookup(tab, nt, data, nfloat, "FF_b", &m->FF_b);
    if (0 != rc) {
        sb2_set_error("weight blob: missing or out-of-range tensor");
        sb2_host_model_free(m);
        return -1;
    }
    if (m->arch == 2) {
        /* events model (src/networks.c:146-194): 12 windowed features -> 2 x (LSTM pair + feedforward2_tanh) -> softmax;
         * sW[l] = [H][4H] recurrent weights, sW2[l] = the 3H peepholes */
        m->H = m->sW[0].nr;
        m->ffw = m->comb_b[0].nr;
        m->nfilter = m->iW[0].nr;
        m->nstate = m->FF_W.nc;
        m->ostride = 4 * ((m->nstate + 3) / 4);
        for (int l = 0; l < 4 && 0 == rc; l++) {
            const uint32_t in = (l < 2) ? m->nfilter : m->ffw;
            if (m->iW[l].nr != in || m->iW[l].stride != in || m->iW[l].nc != 4 * m->H || m->b[l].nr != 4 * m->H ||
                m->sW[l].nr != m->H || m->sW[l].stride != m->H || m->sW[l].nc != 4 * m->H || m->sW2[l].nr != 3 * m->H)
                rc = -1;
        }
        for (int i = 0; i < 2 && 0 == rc; i++)
            if (m->comb_Wf[i].nr != m->H || m->comb_Wb[i].nr != m->H || m->comb_Wf[i].nc != m->ffw ||
                m->comb_Wb[i].nc != m->ffw || m->comb_Wf[i].stride != m->H || m->comb_Wb[i].stride != m->H)
                rc = -1;
        if (m->FF_W.nr != m->ffw || m->FF_W.stride != m->ffw || m->nfilter % 4 != 0) rc = -1;
        if (0 != rc) {
            sb2_set_error("weight blob: unexpected shapes in the events model");
            sb2_host_model_free(m);
            return -1;
        }
        return 0;
    }
    m->winlen = m->conv_W.stride / 4;       /* taps sit at every 4th float of a filter column */
    m->nfilter = m->conv_W.nc;
    m->H = m->sW2[0].nc;
    m->nstate = m->FF_W.nc;
    m->ostride = 4 * ((m->nstate + 3) / 4);
    /* shape sanity: GRU blocks of width H; input widths follow the topology */
    for (int l = 0; l < nlayer; l++) {
        uint32_t in = m->H;
        if (m->arch == 1) in = (l < 2) ? m->nfilter : m->comb_b[0].nr;
        if (m->iW[l].nr != in || m->iW[l].nc != 3 * m->H || m->sW[l].nc != 2 * m->H ||
            m->sW[l].nr != m->H || m->sW2[l].nc != m->H || m->sW2[l].nr != m->H ||
            m->b[l].nr != 3 * m->H || m->iW[l].stride != in) {
            sb2_set_error("weight blob: unexpected GRU shapes in layer %d", l + 1);
            sb2_host_model_free(m);
            return -1;
        }
    }
    if (m->arch == 0) {
        if (m->nfilter != m->H || m->FF_W.nr != m->H || m->H % 4 != 0) { sb2_host_model_free(m); return -1; }
    } else {
        m->ffw = m->comb_b[0].nr;
        for (int i = 0; i < 2; i++)
            if (m->comb_Wf[i].nr != m->H || m->comb_Wb[i].nr != m->H || m->comb_Wf[i].nc != m->ffw ||
                m->comb_Wb[i].nc != m->ffw || m->comb_b[i].nr != m->ffw || m->comb_Wf[i].stride != m->H) {
                sb2_set_error("weight blob: unexpected feedforward2 shapes");
                sb2_host_model_free(m);
                return -1;
            }
        if (m->FF_W.nr != m->ffw || m->FF_W.stride != m->ffw || m->H % 4 != 0 || m->ffw % 4 != 0 || m->nfilter % 4 != 0) {
            sb2_host_model_free(m);
            return -1;
        }
    }
    return 0;
}

void sb2_host_model_free(sb2_host_model *m) {
    if (NULL == m) return;
    free(m->blob);
    memset(m, 0, sizeof(*m));
}

int sb2_read_file(const char *path, void **data, size_t *nbytes) {
    FILE *fh = fopen(path, "rb");
    if (NULL == fh) { sb2_set_error("cannot open %s", path); return -1; }
    fseek(fh, 0, SEEK_END);
    const long sz = ftell(fh);
    fseek(fh, 0, SEEK_SET);
    void *buf = (sz > 0) ? malloc((size_t)sz) : NULL;
    if (NULL == buf || fread(buf, 1, (size_t)sz, fh) != (size_t)sz) {
        free(buf);
        fclose(fh);
        sb2_set_error("cannot read %s", path);
        return -1;
    }
    fclose(fh);
    *data = buf;
    *nbytes = (size_t)sz;
    return 0;
}

int sb2_default_weights_dir(char *buf, size_t buflen) {
    const char *env = getenv("SCRAPPIE_B200_WEIGHTS");
    if (NULL != env && env[0] != '\0') {
        snprintf(buf, buflen, "%s", env);
        return 0;
    }
    Dl_info info;
    if (0 == dladdr((void *)&sb2_default_weights_dir, &info) || NULL == info.dli_fname) return -1;
    snprintf(buf, buflen, "%s", info.dli_fname);
    char *slash = strrchr(buf, '/');
    if (NULL == slash) snprintf(buf, buflen, "weights");
    else snprintf(slash + 1, buflen - (size_t)(slash + 1 - buf), "weights");
    return 0;
}

/* -------------------------------------------------------- convolution plan */

static int plan_add(sb2_conv_tail *plan, long col, long x0, long tap0, long ntap) {
    if (col < plan->first_col || col >= plan->ncol || ntap <= 0) return 0;
    const int c = (int)(col - plan->first_col);
    if (plan->nseg[c] >= SB2_CONV_TAIL_SEGS) return -1;
    int32_t *s = plan->seg[c][plan->nseg[c]++];
    s[0] = (int32_t)x0; s[1] = (int32_t)tap0; s[2] = (int32_t)ntap;
    return 0;
}

/* Which products the reference's convolution() accumulates into the last
 * SB2_CONV_TAIL_COLS output columns, derived from its index arithmetic
 * (src/layers.c:190-241): left-edge calls, the strided body (one call per offset
 * w = 0, stride, ... < winlen, each covering floor((n - shift - w) / nstepX) windows)
 * and the right-edge calls (placed at offsetC_R + w / stride).  Columns before
 * first_col are plain zero-padded same-convolution windows; the CUDA kernel computes
 * those arithmetically. */
int sb2_conv_plan(size_t nsample, size_t winlen, size_t stride, sb2_conv_tail *plan) {
    memset(plan, 0, sizeof(*plan));
    if (0 == stride || winlen < 1 || nsample < winlen) {
        sb2_set_error("convolution: read of %zu samples is shorter than the %zu-sample window", nsample, winlen);
        return -1;
    }
    const long n = (long)nsample, W = (long)winlen, S = (long)stride;
    const long padL = (W - 1) / 2, padR = W / 2;
    const long ncol = (n + S - 1) / S;
    plan->ncol = (int32_t)ncol;
    plan->first_col = (int32_t)(ncol > SB2_CONV_TAIL_COLS ? ncol - SB2_CONV_TAIL_COLS : 0);
    int rc = 0;

    for (long w = 0; w < padL; w += S) rc |= plan_add(plan, w / S, 0, padL - w, W - (padL - w));

    const long ncolL = (padL + S - 1) / S;
    const long shift = ncolL * S - padL;
    const long nstepC = (W + S - 1) / S;
    const long nstepX = S * nstepC;
    for (long w = 0; w < W; w += S) {
        const long nwin = (n - shift - w > 0) ? (n - shift - w) / nstepX : 0;
        /* only windows that can reach the tail need enumerating */
        long j0 = (plan->first_col - (w / S + ncolL)) / nstepC - 1;
        if (j0 < 0) j0 = 0;
        for (long j = j0; j < nwin; j++)
            rc |= plan_add(plan, w / S + ncolL + j * nstepC, shift + w + j * nstepX, 0, W);
    }

    const long maxcol = (n - shift) / nstepX;
    const long rem = (n - shift) % nstepX;
    const long colR = ncolL + nstepC * (maxcol - 1) + rem / S + 1;
    const long xR = n - W + 1;
    const long startR = S - (padL + n - W) % S - 1;
    for (long w = startR; w < padR; w += S) rc |= plan_add(plan, colR + w / S, xR + w, 0, W - 1 - w);

    if (0 != rc) sb2_set_error("convolution plan: too many segments in one column");
    return rc;
}

/* Flat dump of the plan for tests: out = {first_col, ncol, then per tail column
 * {nseg, (x0, tap0, ntap) x 3}}. */
int sb2_conv_plan_debug(size_t nsample, size_t winlen, size_t stride, int *out, int nout) {
    sb2_conv_tail plan;
    if (NULL == out || nout < 2 + SB2_CONV_TAIL_COLS * (1 + 3 * SB2_CONV_TAIL_SEGS)) return -1;
    if (0 != sb2_conv_plan(nsample, winlen, stride, &plan)) return -1;
    out[0] = plan.first_col;
    out[1] = plan.ncol;
    int *p = out + 2;
    for (int c = 0; c < SB2_CONV_TAIL_COLS; c++) {
        *p++ = plan.nseg[c];
        for (int s = 0; s < SB2_CONV_TAIL_SEGS; s++)
            for (int k = 0; k < 3; k++) *p++ = plan.seg[c][s][k];
    }
    return 0;
}
