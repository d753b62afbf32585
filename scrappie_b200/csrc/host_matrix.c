/* Host containers of the libscrappie ABI (scrappie_matrix).
 *
 * Behavioural mirror of src/scrappie_matrix.c:11-136 in the reference: column-major
 * fp32, columns padded to a multiple of four floats, 16-byte aligned, zero filled,
 * NULL on any allocation failure, free returns NULL.
 */
#include <stdlib.h>
#include <string.h>

#include "scrappie_b200.h"

scrappie_matrix make_scrappie_matrix(size_t nr, size_t nc) {
    if (0 == nr || 0 == nc) return NULL;
    const size_t nrq = (nr + 3) / 4;
    const size_t colbytes = nrq * 4 * sizeof(float);
    if (colbytes / (4 * sizeof(float)) != nrq) return NULL;
    const size_t total = colbytes * nc;
    if (total / colbytes != nc) return NULL;             /* size overflow */

    scrappie_matrix m = malloc(sizeof(*m));
    if (NULL == m) return NULL;
    void *buf = NULL;
    if (0 != posix_memalign(&buf, 16, total)) {
        free(m);
        return NULL;
    }
    memset(buf, 0, total);
    m->nr = nr;
    m->nrq = nrq;
    m->nc = nc;
    m->stride = nrq * 4;
    m->data.v = buf;
    return m;
}

scrappie_matrix remake_scrappie_matrix(scrappie_matrix M, size_t nr, size_t nc) {
    if (NULL != M && M->nr == nr && M->nc == nc) return M;
    free_scrappie_matrix(M);
    return make_scrappie_matrix(nr, nc);
}

scrappie_matrix copy_scrappie_matrix(const_scrappie_matrix M) {
    if (NULL == M) return NULL;
    scrappie_matrix C = make_scrappie_matrix(M->nr, M->nc);
    if (NULL == C) return NULL;
    memcpy(C->data.f, M->data.f, M->stride * M->nc * sizeof(float));
    return C;
}

scrappie_matrix free_scrappie_matrix(scrappie_matrix mat) {
    if (NULL != mat) {
        free(mat->data.v);
        free(mat);
    }
    return NULL;
}

void zero_scrappie_matrix(scrappie_matrix M) {
    if (NULL == M) return;
    memset(M->data.f, 0, M->stride * M->nc * sizeof(float));
}

scrappie_matrix mat_from_array(const float *x, size_t nr, size_t nc) {
    if (NULL == x) return NULL;
    scrappie_matrix m = make_scrappie_matrix(nr, nc);
    if (NULL == m) return NULL;
    for (size_t c = 0; c < nc; c++) memcpy(m->data.f + c * m->stride, x + c * nr, nr * sizeof(float));
    return m;
}

float *array_from_scrappie_matrix(const_scrappie_matrix mat) {
    if (NULL == mat) return NULL;
    float *out = calloc(mat->nr * mat->nc, sizeof(float));
    if (NULL == out) return NULL;
    for (size_t c = 0; c < mat->nc; c++)
        memcpy(out + c * mat->nr, mat->data.f + c * mat->stride, mat->nr * sizeof(float));
    return out;
}
