/* Minimal C caller of libscrappie_b200.so -- the shape of `scrappie raw` with the GPU engine behind it.
 *
 *   gcc -std=c99 -Iinclude examples/raw_basecall.c -Lscrappie_b200 -lscrappie_b200 -Wl,-rpath,$PWD/scrappie_b200 -lm
 *   ./a.out rgrgr_r94 signal.f32 [more.f32 ...]        (files of raw pA samples as little-endian float32)
 *
 * With no file arguments it only exercises the host side (registry, containers, signal preparation), which needs
 * no GPU; that mode is what tests/test_host.py compiles and runs.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "scrappie_b200.h"

static float *read_f32(const char *path, size_t *n) {
    FILE *fh = fopen(path, "rb");
    if (NULL == fh) return NULL;
    fseek(fh, 0, SEEK_END);
    const long sz = ftell(fh);
    fseek(fh, 0, SEEK_SET);
    float *x = (sz > 0) ? malloc((size_t)sz) : NULL;
    if (NULL != x && fread(x, 1, (size_t)sz, fh) != (size_t)sz) { free(x); x = NULL; }
    fclose(fh);
    *n = (NULL != x) ? (size_t)sz / sizeof(float) : 0;
    return x;
}

int main(int argc, char **argv) {
    const char *model_name = (argc > 1) ? argv[1] : "rgrgr_r94";
    const enum raw_model_type model = get_raw_model(model_name);
    if (SCRAPPIE_MODEL_INVALID == model) { fprintf(stderr, "unknown model %s\n", model_name); return EXIT_FAILURE; }
    printf("model %s stride %d\n", raw_model_string(model), get_raw_model_stride(model));

    if (argc <= 2) {                                   /* host-only self check */
        float x[400];
        for (int i = 0; i < 400; i++) x[i] = 90.0f + 12.0f * sinf(0.37f * (float)i) + (float)(i % 7);
        raw_table rt = {NULL, 400, 0, 400, malloc(sizeof(x))};
        memcpy(rt.raw, x, sizeof(x));
        rt = trim_and_segment_raw(rt, 200, 10, 100, 0.0f);
        if (NULL == rt.raw) return EXIT_FAILURE;
        medmad_normalise_array(rt.raw + rt.start, rt.end - rt.start);
        scrappie_matrix m = mat_from_array(rt.raw + rt.start, 1, rt.end - rt.start);
        printf("trimmed to [%zu, %zu), matrix %zu x %zu stride %zu, median now %.3f\n", rt.start, rt.end, m->nr, m->nc,
               m->stride, medianf(rt.raw + rt.start, rt.end - rt.start));
        m = free_scrappie_matrix(m);
        free(rt.raw);
        return (NULL == m) ? EXIT_SUCCESS : EXIT_FAILURE;
    }

    /* calculate_post (src/scrappie_raw.c:265-315) for all the files at once, everything on the device */
    const size_t nread = (size_t)argc - 2;
    const float **raw = calloc(nread, sizeof(*raw));
    size_t *n = calloc(nread, sizeof(*n)), *start = calloc(nread, sizeof(*start)), *end = calloc(nread, sizeof(*end));
    sb2_call *calls = calloc(nread, sizeof(*calls));
    for (size_t i = 0; i < nread; i++) raw[i] = read_f32(argv[i + 2], &n[i]);
    sb2_engine *eng = sb2_engine_create(0, NULL);
    if (NULL == eng) { fprintf(stderr, "%s\n", sb2_last_error()); return EXIT_FAILURE; }
    const sb2_params p = sb2_default_params();
    const sb2_trim t = sb2_default_trim();
    const int ncalled = sb2_basecall_raw_batch(eng, model, raw, n, nread, &t, &p, calls, start, end);
    for (size_t i = 0; i < nread; i++) {
        if (NULL == calls[i].bases) continue;
        /* the FASTA record of src/scrappie_raw.c:317-331 */
        printf(">%s  { \"normalised_score\" : %f,  \"nblock\" : %zu,  \"sequence_length\" : %zu,  \"blocks_per_base\" : %f, "
               "\"nsample\" : %zu, \"trim\" : [ %zu, %zu ] }\n%s\n", argv[i + 2],
               -calls[i].score / (float)calls[i].nblock, calls[i].nblock, calls[i].nbase,
               (float)calls[i].nblock / (float)calls[i].nbase, n[i], start[i], end[i], calls[i].bases);
    }
    sb2_calls_free(calls, nread);
    sb2_engine_destroy(eng);
    for (size_t i = 0; i < nread; i++) free((void *)raw[i]);
    free(calls); free(end); free(start); free(n); free(raw);
    return (ncalled >= 0) ? EXIT_SUCCESS : EXIT_FAILURE;
}
