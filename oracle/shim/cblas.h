/* Test infrastructure only (oracle build): minimal CBLAS declarations so the
 * reference's C sources can be compiled where they lie under /root/reference
 * against the OpenBLAS that scipy bundles in this image (symbols are prefixed
 * `scipy_`).  Only the two BLAS entry points the reference calls are declared
 * (call sites: src/layers.c:193,220,237,505,517 and src/scrappie_matrix.c:346). */
#ifndef SB2_ORACLE_CBLAS_SHIM_H
#define SB2_ORACLE_CBLAS_SHIM_H

enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };

#define cblas_sgemm scipy_cblas_sgemm
#define cblas_sgemv scipy_cblas_sgemv
#define openblas_set_num_threads scipy_openblas_set_num_threads

void cblas_sgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb,
                 int M, int N, int K, float alpha, const float *A, int lda,
                 const float *B, int ldb, float beta, float *C, int ldc);
void cblas_sgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, int M, int N,
                 float alpha, const float *A, int lda, const float *X, int incX,
                 float beta, float *Y, int incY);
void openblas_set_num_threads(int n);

#endif
