/* O(T) integer post-processing of a Viterbi path on the host (stays on the CPU, as in
 * the reference): k-mer path -> bases, CRF path -> bases, homopolymer length fix-up.
 *
 * Behavioural mirrors of src/decode.c:367-382 (overlap), :449-509 (overlapper),
 * :895-918 (crfpath_to_basecall) and src/homopolymer.c:67-235 (findRuns,
 * homopolymer_path).  Error convention of the reference: NULL / -1, never abort.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "scrappie_b200.h"
#include "sb2_internal.h"

static const char base_of[4] = {'A', 'C', 'G', 'T'};

/* Number of new bases when k-mer `next` follows k-mer `prev`: the smallest shift s >= 1
 * for which the last (k-s) bases of prev equal the first (k-s) bases of next. */
static int kmer_shift(int prev, int next, int nkmer) {
    int mask = nkmer - 1;
    int shift = 0;
    do {
        mask >>= 2;
        prev &= mask;
        next >>= 2;
        shift++;
    } while (prev != next);
    return shift;
}

static size_t kmer_length_of(int nkmer) {
    size_t bits = 0;
    for (size_t x = (size_t)nkmer; x != 0; x >>= 1) bits++;
    return bits / 2;
}

char *overlapper(const int *seq, size_t n, int nkmer, int *pos) {
    if (NULL == seq) return NULL;
    const size_t klen = kmer_length_of(nkmer);

    size_t first = 0;
    while (first < n && seq[first] < 0) first++;
    if (first == n) return NULL;                /* all stays: nothing to call */

    size_t nbase = klen;
    for (size_t i = first + 1, prev = first; i < n; i++) {
        if (seq[i] < 0) continue;
        nbase += (size_t)kmer_shift(seq[prev], seq[i], nkmer);
        prev = i;
    }
    char *bases = calloc(nbase + 1, sizeof(char));
    if (NULL == bases) return NULL;

    for (size_t j = 0, kmer = (size_t)seq[first]; j < klen; j++, kmer >>= 2)
        bases[klen - 1 - j] = base_of[kmer & 3];
    if (NULL != pos) pos[0] = 0;

    size_t tail = klen - 1;                     /* index of the last base written */
    int prev = seq[first];
    for (size_t i = first + 1; i < n; i++) {
        if (seq[i] < 0) {
            if (NULL != pos) pos[i] = pos[i - 1];
            continue;
        }
        const int shift = kmer_shift(prev, seq[i], nkmer);
        if (NULL != pos) pos[i] = pos[i - 1] + shift;
        size_t kmer = (size_t)seq[i];
        for (int j = 0; j < shift; j++, kmer >>= 2) bases[tail + shift - j] = base_of[kmer & 3];
        tail += shift;
        prev = seq[i];
    }
    return bases;
}

char *crfpath_to_basecall(int const *path, size_t npos, int *pos) {
    /* the reference requires a non-NULL pos but never writes to it (:895-918) */
    if (NULL == path || NULL == pos) return NULL;
    size_t nbase = 0;
    for (size_t i = 0; i < npos; i++) nbase += (path[i] < 4);
    char *bases = calloc(nbase + 1, sizeof(char));
    if (NULL == bases) return NULL;
    for (size_t i = 0, j = 0; i < npos; i++)
        if (path[i] < 4) bases[j++] = base_of[path[i]];
    return bases;
}

enum homopolymer_calculation get_homopolymer_calculation(const char *calcstr) {
    if (NULL == calcstr) return HOMOPOLYMER_INVALID;
    if (0 == strcmp(calcstr, "nochange")) return HOMOPOLYMER_NOCHANGE;
    if (0 == strcmp(calcstr, "mean")) return HOMOPOLYMER_MEAN;
    return HOMOPOLYMER_INVALID;
}

static int homopolymer_kmer(int base, int len) {
    int k = 0;
    for (int i = 0; i < len; i++) k = 4 * k + base;
    return k;
}

int sb2_find_homopolymer_runs(const int *path, int pathlen, int klen, sb2_hp_run **runs_out) {
    *runs_out = NULL;
    if (NULL == path) return -1;
    const int cap = pathlen / 2 > 0 ? pathlen / 2 : 1;
    sb2_hp_run *runs = calloc((size_t)cap, sizeof(*runs));
    if (NULL == runs) return -1;
    const int mod1 = 1 << (2 * (klen - 1)), mod2 = 1 << (2 * (klen - 2));
    int n = 0;
    /* Scan order matters (runs are applied sequentially and may overlap): base-major,
     * then position, first the "XYYYY" rule then the "ZXYYY" rule -- homopolymer.c:94-137 */
    for (int base = 0; base < 4; base++) {
        const int full = homopolymer_kmer(base, klen);
        const int tail1 = homopolymer_kmer(base, klen - 1);
        const int tail2 = homopolymer_kmer(base, klen - 2);
        for (int i = 1; i < pathlen - 2; i++) {
            const int before = path[i - 1], here = path[i];
            const int here_ok = (here == -1) || (here == full);
            if (before == -1 || !here_ok) continue;
            if ((before % mod1 == tail1) && before != full) {
                int e = i + 1;
                while (e < pathlen && (path[e] == -1 || path[e] == full)) e++;
                runs[n].start = i; runs[n].length = e - i; runs[n].state = full; n++;
            }
            if ((before % mod2 == tail2) && (before % mod1 != tail1)) {
                int j = i;
                while (j < pathlen && path[j] == -1) j++;
                if (path[j] == full && j < pathlen - 1) {
                    int e = j + 1;
                    while (e < pathlen && (path[e] == -1 || path[e] == full)) e++;
                    runs[n].start = j; runs[n].length = e - j; runs[n].state = full; n++;
                }
            }
        }
    }
    *runs_out = runs;
    return n;
}

void sb2_apply_homopolymer_run(int *path, const sb2_hp_run *run, const float *logp_stay,
                               const float *logp_rep) {
    int nviterbi = 0;
    double expect = 0.0;
    for (int i = 0; i < run->length; i++) {
        const double ps = expf(logp_stay[i]);
        const double pr = expf(logp_rep[i]);
        expect += pr / (pr + ps);
        if (path[run->start + i] == run->state) nviterbi++;
    }
    const int nnew = (int)(expect + 0.5);
    if (nnew == nviterbi) return;
    for (int i = 0; i < run->length; i++) path[run->start + i] = (i < nnew) ? run->state : -1;
}

int homopolymer_path(const_scrappie_matrix post, int *viterbipath,
                     enum homopolymer_calculation pathCalculationFlag) {
    if (pathCalculationFlag != HOMOPOLYMER_MEAN) return 0;
    if (NULL == post || NULL == viterbipath) return -1;
    const int nblock = (int)post->nc;
    const int stay = (int)post->nr - 1;
    const int klen = (int)(logf((float)post->nr) / logf(4.0f));
    sb2_hp_run *runs = NULL;
    const int nrun = sb2_find_homopolymer_runs(viterbipath, nblock, klen, &runs);
    if (nrun < 0) return nrun;
    int maxlen = 1;
    for (int r = 0; r < nrun; r++) if (runs[r].length > maxlen) maxlen = runs[r].length;
    float *ps = malloc(2 * (size_t)maxlen * sizeof(float));
    if (NULL == ps) { free(runs); return -1; }
    float *pr = ps + maxlen;
    for (int r = 0; r < nrun; r++) {
        /* path[i] pairs with posterior column i-1 (homopolymer.c:205-217) */
        for (int i = 0; i < runs[r].length; i++) {
            const size_t col = (size_t)(runs[r].start + i - 1);
            ps[i] = post->data.f[col * post->stride + stay];
            pr[i] = post->data.f[col * post->stride + runs[r].state];
        }
        sb2_apply_homopolymer_run(viterbipath, &runs[r], ps, pr);
    }
    free(ps);
    free(runs);
    return 0;
}

/* encode_bases_to_integers, src/scrappie_seq_helpers.c:15-75: A/C/G/T (either case) -> base-4 k-mer states;
 * NULL when a base is not recognised (the reference also warns on stderr). */
int *encode_bases_to_integers(char const *seq, size_t n, size_t state_len) {
    if (NULL == seq || 0 == state_len || n < state_len) return NULL;
    const size_t nstate = n - state_len + 1;
    int *iseq = calloc(nstate, sizeof(int));
    if (NULL == iseq) return NULL;
    for (size_t i = 0; i < nstate; i++) {
        int ib = 0;
        for (size_t j = 0; j < state_len; j++) {
            int b;
            switch (seq[i + j]) {
            case 'A': case 'a': b = 0; break;
            case 'C': case 'c': b = 1; break;
            case 'G': case 'g': b = 2; break;
            case 'T': case 't': b = 3; break;
            default: free(iseq); return NULL;
            }
            ib = ib * 4 + b;
        }
        iseq[i] = ib;
    }
    return iseq;
}
