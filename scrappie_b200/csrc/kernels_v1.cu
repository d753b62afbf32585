// First-generation CUDA kernels of the raw basecalling path (fp32 CUDA-core arithmetic).
// They are the correctness baseline for the tcgen05 kernels in kernels_tc.cu and stay
// selectable at run time (SCRAPPIE_B200_SCAN=ffma, SCRAPPIE_B200_GEMM=ffma).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "device_math.cuh"
#include "kernels.h"

namespace sb2 {

// ---------------------------------------------------------------------------------
// convolution + activation
// ---------------------------------------------------------------------------------
// One CTA = CONV_CPB consecutive output columns of one read.  Taps are staged in shared
// memory transposed to [tap][filter] so that a warp reads consecutive filters; the
// samples a CTA needs are staged once.  Columns in the read's tail follow the explicit
// plan (see sb2_conv_plan), everything else is a zero padded same-convolution window.
constexpr int CONV_CPB = 64;

__global__ void __launch_bounds__(256)
conv_act_kernel(const float *__restrict__ raw, BatchDims d, const sb2_conv_tail *__restrict__ tails,
                const float *__restrict__ taps, const float *__restrict__ bias, int winlen, int nf,
                int nfp, int stride, int act, float *__restrict__ out) {
    extern __shared__ float smem[];
    const int r = blockIdx.y;
    const int ncol = d.nblock[r];
    const int c0 = blockIdx.x * CONV_CPB;
    if (c0 >= ncol) return;
    const int n = d.nsample[r];
    const float *x = raw + d.samp_off[r];
    const sb2_conv_tail *tail = tails + r;
    const int padL = (winlen - 1) / 2;

    float *s_taps = smem;                               // [winlen][nf]
    float *s_x = smem + winlen * nf;                    // samples [xlo, xlo + nx)
    const int nx = (CONV_CPB - 1) * stride + winlen;
    const int xlo = c0 * stride - padL;
    for (int i = threadIdx.x; i < winlen * nf; i += blockDim.x) s_taps[i] = taps[i];
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
        const int xi = xlo + i;
        s_x[i] = (xi >= 0 && xi < n) ? x[xi] : 0.0f;
    }
    __syncthreads();

    const int f = threadIdx.x % nfp;
    const int lane = threadIdx.x / nfp;
    const int nlane = blockDim.x / nfp;
    if (f >= nf || lane >= nlane) return;
    const float bf = bias[f];
    const int first_tail = tail->first_col;
    for (int cc = lane; cc < CONV_CPB; cc += nlane) {
        const int c = c0 + cc;
        if (c >= ncol) break;
        float acc = 0.0f;
        if (c < first_tail) {
            int x0 = c * stride - padL, tap0 = 0;
            if (x0 < 0) { tap0 = -x0; x0 = 0; }
            for (int k = tap0; k < winlen; k++) acc = fmaf(s_taps[k * nf + f], s_x[x0 - xlo + (k - tap0)], acc);
        } else {
            const int tc = c - first_tail;
            const int nseg = tail->nseg[tc];
            for (int sg = 0; sg < nseg; sg++) {
                const int x0 = tail->seg[tc][sg][0], tap0 = tail->seg[tc][sg][1], ntap = tail->seg[tc][sg][2];
                float part = 0.0f;
                for (int k = 0; k < ntap; k++) part = fmaf(s_taps[(tap0 + k) * nf + f], x[x0 + k], part);
                acc += part;
            }
        }
        float v = bf + acc;
        v = (act == 0) ? elu_cephes(v) : tanh_cephes(v);
        out[(size_t)(d.col_off[r] + c) * nf + f] = v;
    }
}

// Second generation: one thread = one filter, its taps live in registers, and it produces four
// consecutive columns from one shared-memory window (3 * STRIDE + WINLEN samples read once, broadcast
// within the warp) -- 4 * WINLEN FMAs per window instead of two shared-memory loads per FMA.
// Columns that touch the left edge or the planned tail take the generic per-column path.
template <int WINLEN, int STRIDE>
__global__ void __launch_bounds__(256)
conv_act_v2_kernel(const float *__restrict__ raw, BatchDims d, const sb2_conv_tail *__restrict__ tails,
                   const float *__restrict__ taps, const float *__restrict__ bias, int nf, int nfp, int act,
                   float *__restrict__ out) {
    constexpr int CPB = CONV_CPB;
    constexpr int NX = (CPB - 1) * STRIDE + WINLEN;
    constexpr int PADL = (WINLEN - 1) / 2;
    constexpr int WIN4 = 3 * STRIDE + WINLEN;
    __shared__ float s_x[NX];
    const int r = blockIdx.y;
    const int ncol = d.nblock[r];
    const int c0 = blockIdx.x * CPB;
    if (c0 >= ncol) return;
    const int n = d.nsample[r];
    const float *x = raw + d.samp_off[r];
    const sb2_conv_tail *tail = tails + r;
    const int xlo = c0 * STRIDE - PADL;
    for (int i = threadIdx.x; i < NX; i += blockDim.x) {
        const int xi = xlo + i;
        s_x[i] = (xi >= 0 && xi < n) ? x[xi] : 0.0f;
    }
    const int f = threadIdx.x % nfp;
    const int lane = threadIdx.x / nfp;
    const int nlane = blockDim.x / nfp;
    float w[WINLEN];
    const bool active = (f < nf) && (lane < nlane);
#pragma unroll
    for (int k = 0; k < WINLEN; k++) w[k] = active ? taps[k * nf + f] : 0.0f;
    const float bf = active ? bias[f] : 0.0f;
    __syncthreads();
    if (!active) return;
    const int first_tail = tail->first_col;
    float *orow = out + (size_t)d.col_off[r] * nf + f;
    for (int cc = 4 * lane; cc < CPB; cc += 4 * nlane) {
        const int c = c0 + cc;
        if (c >= ncol) break;
        if (c * STRIDE - PADL >= 0 && c + 3 < first_tail && c + 3 < ncol) {
            float xs[WIN4];
#pragma unroll
            for (int i = 0; i < WIN4; i++) xs[i] = s_x[cc * STRIDE + i];
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < WINLEN; k++) {
#pragma unroll
                for (int j = 0; j < 4; j++) acc[j] = fmaf(w[k], xs[j * STRIDE + k], acc[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float v = bf + acc[j];
                v = (act == 0) ? elu_cephes(v) : tanh_cephes(v);
                orow[(size_t)(c + j) * nf] = v;
            }
        } else {
            for (int j = 0; j < 4 && c + j < ncol; j++) {
                const int cj = c + j;
                float acc = 0.0f;
                if (cj < first_tail) {
                    int x0 = cj * STRIDE - PADL, tap0 = 0;
                    if (x0 < 0) { tap0 = -x0; x0 = 0; }
                    for (int k = tap0; k < WINLEN; k++) acc = fmaf(taps[k * nf + f], s_x[x0 - xlo + (k - tap0)], acc);
                } else {
                    const int tc = cj - first_tail;
                    const int nseg = tail->nseg[tc];
                    for (int sg = 0; sg < nseg; sg++) {
                        const int x0 = tail->seg[tc][sg][0], tap0 = tail->seg[tc][sg][1], ntap = tail->seg[tc][sg][2];
                        float part = 0.0f;
                        for (int k = 0; k < ntap; k++) part = fmaf(taps[(tap0 + k) * nf + f], x[x0 + k], part);
                        acc += part;
                    }
                }
                float v = bf + acc;
                v = (act == 0) ? elu_cephes(v) : tanh_cephes(v);
                orow[(size_t)cj * nf] = v;
            }
        }
    }
}

void launch_conv_act(const float *raw, const BatchDims &d, const sb2_conv_tail *tails, const float *taps,
                     const float *bias, int winlen, int nf, int stride, int act, float *out,
                     cudaStream_t s) {
    const int nfp = (nf + 31) / 32 * 32;
    dim3 grid((d.max_cols + CONV_CPB - 1) / CONV_CPB, d.nread);
    // whole column lanes only: 96 filters -> 192 threads (two lanes), not 256 with a quarter of the CTA idle
    const int nthr = (nfp <= 256) ? nfp * (256 / nfp) : 256;
    if (winlen == 19 && stride == 5) {
        conv_act_v2_kernel<19, 5><<<grid, nthr, 0, s>>>(raw, d, tails, taps, bias, nf, nfp, act, out);
        return;
    }
    if (winlen == 11 && stride == 1) {
        conv_act_v2_kernel<11, 1><<<grid, nthr, 0, s>>>(raw, d, tails, taps, bias, nf, nfp, act, out);
        return;
    }
    if (winlen == 11 && stride == 5) {
        conv_act_v2_kernel<11, 5><<<grid, nthr, 0, s>>>(raw, d, tails, taps, bias, nf, nfp, act, out);
        return;
    }
    const size_t smem = (size_t)(winlen * nf + (CONV_CPB - 1) * stride + winlen) * sizeof(float);
    conv_act_kernel<<<grid, 256, smem, s>>>(raw, d, tails, taps, bias, winlen, nf, nfp, stride, act, out);
}

// ---------------------------------------------------------------------------------
// affine map  C = f((b + W^T X) / cdiv)   (fp32 register-tiled GEMM)
// ---------------------------------------------------------------------------------
constexpr int AF_BM = 96;       // output rows per CTA
constexpr int AF_BN = 128;      // columns per CTA
constexpr int AF_KC = 32;       // K chunk
constexpr int AF_XS = AF_BN + 4;

__global__ void __launch_bounds__(256)
affine_kernel(const float *__restrict__ X, int ncol, int K, const float *__restrict__ W, int ldw,
              const float *__restrict__ b, int M, float *__restrict__ C, int ldc, float xdiv,
              float cdiv, int act, int accumulate) {
    __shared__ __align__(16) float Ws[AF_KC][AF_BM];
    __shared__ __align__(16) float Xs[AF_KC][AF_XS];
    const int m_base = blockIdx.y * AF_BM;
    const int n_base = blockIdx.x * AF_BN;
    const int tm = threadIdx.x % 16, tn = threadIdx.x / 16;
    const int m0 = tm * 6, n0 = tn * 8;
    float acc[6][8];
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.0f;

    for (int k0 = 0; k0 < K; k0 += AF_KC) {
        // W chunk -> Ws[k][m]   (consecutive threads take consecutive rows: conflict-free stores)
        for (int idx = threadIdx.x; idx < AF_BM * (AF_KC / 4); idx += 256) {
            const int m = idx % AF_BM, k4 = idx / AF_BM;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m_base + m < M && k0 + 4 * k4 < K)
                v = *reinterpret_cast<const float4 *>(W + (size_t)(m_base + m) * ldw + k0 + 4 * k4);
            Ws[4 * k4 + 0][m] = v.x; Ws[4 * k4 + 1][m] = v.y; Ws[4 * k4 + 2][m] = v.z; Ws[4 * k4 + 3][m] = v.w;
        }
        for (int idx = threadIdx.x; idx < AF_BN * (AF_KC / 4); idx += 256) {
            const int c = idx % AF_BN, k4 = idx / AF_BN;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n_base + c < ncol && k0 + 4 * k4 < K)
                v = *reinterpret_cast<const float4 *>(X + (size_t)(n_base + c) * K + k0 + 4 * k4);
            Xs[4 * k4 + 0][c] = v.x / xdiv; Xs[4 * k4 + 1][c] = v.y / xdiv;
            Xs[4 * k4 + 2][c] = v.z / xdiv; Xs[4 * k4 + 3][c] = v.w / xdiv;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < AF_KC; k++) {
            float a[6], bb[8];
            const float2 a0 = *reinterpret_cast<const float2 *>(&Ws[k][m0]);
            const float2 a1 = *reinterpret_cast<const float2 *>(&Ws[k][m0 + 2]);
            const float2 a2 = *reinterpret_cast<const float2 *>(&Ws[k][m0 + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a1.x; a[3] = a1.y; a[4] = a2.x; a[5] = a2.y;
            const float4 b0 = *reinterpret_cast<const float4 *>(&Xs[k][n0]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Xs[k][n0 + 4]);
            bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w;
            bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
#pragma unroll
            for (int i = 0; i < 6; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int c = n_base + n0 + j;
        if (c >= ncol) continue;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const int m = m_base + m0 + i;
            if (m >= M) continue;
            // accumulate: second half of affine_map2 (src/scrappie_matrix.c:353-383), C already holds b + W1 x1
            const float base = accumulate ? C[(size_t)c * ldc + m] : b[m];
            float v = (base + acc[i][j]) / cdiv;
            if (act == 1) v = exp_cephes(v);
            else if (act == 2) v = tanh_cephes(v);
            C[(size_t)c * ldc + m] = v;
        }
    }
}

void launch_affine(const float *X, int ncol, int K, const float *W, int ldw, const float *b, int M,
                   float *C, int ldc, float xdiv, float cdiv, int act, int accumulate, cudaStream_t s) {
    dim3 grid((ncol + AF_BN - 1) / AF_BN, (M + AF_BM - 1) / AF_BM);
    affine_kernel<<<grid, 256, 0, s>>>(X, ncol, K, W, ldw, b, M, C, ldc, xdiv, cdiv, act, accumulate);
}

// ---------------------------------------------------------------------------------
// softmax normalisation + robust log, one warp per column
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_finish_kernel(float *__restrict__ post, int ncol, int nstate, int ostride, float min_prob,
                      int return_log) {
    const int col = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (col >= ncol) return;
    float *p = post + (size_t)col * ostride;
    float sum = 0.0f;
    for (int k = lane; k < nstate; k += 32) sum += p[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float recip = __fdiv_rn(1.0f, sum);
    const float keep = __fsub_rn(1.0f, min_prob);
    for (int k = lane; k < ostride; k += 32) {
        // padding lanes hold exp(0) = 1 in the reference and are scaled like the rest
        float v = __fmul_rn((k < nstate) ? p[k] : 1.0f, recip);
        if (return_log) v = log_cephes(__fadd_rn(min_prob, __fmul_rn(keep, v)));
        p[k] = v;
    }
}

void launch_softmax_finish(float *post, int ncol, int nstate, int ostride, float min_prob,
                           int return_log, cudaStream_t s) {
    softmax_finish_kernel<<<(ncol + 7) / 8, 256, 0, s>>>(post, ncol, nstate, ostride, min_prob, return_log);
}

// ---------------------------------------------------------------------------------
// GRU scan, fp32 CUDA cores: weights live in registers for the whole layer
// ---------------------------------------------------------------------------------
// One CTA = GRU_R reads stepping together; 3H threads.  Thread k < 2H owns row k of sW
// (update / reset gates), thread 2H + j owns row j of sW2 (candidate).  Per step:
//   phase 1  g = sigma(x[0:2H] + sW^T h);  z -> smem, r*h -> smem
//   phase 2  c = tanh(x[2H:3H] + sW2^T (r*h));  h' = z*h + (1-z)*c
// The hidden state is double buffered in shared memory as [H][GRU_R] so that a thread
// fetches the state of all reads for one input index with two 16-byte broadcasts.
constexpr int GRU_R = 8;

template <int H>
__global__ void __launch_bounds__(3 * H, 1)
gru_scan_ffma_kernel(const float *__restrict__ Xin, const float *__restrict__ sW,
                     const float *__restrict__ sW2, const float *__restrict__ resid,
                     float *__restrict__ out, BatchDims d, int backward) {
    __shared__ __align__(16) float hs[2][H][GRU_R];
    __shared__ __align__(16) float rhs[H][GRU_R];
    __shared__ __align__(16) float zs[H][GRU_R];
    __shared__ int s_T[GRU_R], s_col[GRU_R];

    const int k = threadIdx.x;
    const int r0 = blockIdx.x * GRU_R;
    if (k < GRU_R) {
        const int r = r0 + k;
        s_T[k] = (r < d.nread) ? d.nblock[r] : 0;
        s_col[k] = (r < d.nread) ? d.col_off[r] : 0;
    }
    for (int i = k; i < 2 * H * GRU_R; i += 3 * H) (&hs[0][0][0])[i] = 0.0f;

    float w[H];
    {
        const float *row = (k < 2 * H) ? (sW + (size_t)k * H) : (sW2 + (size_t)(k - 2 * H) * H);
#pragma unroll
        for (int i = 0; i < H; i += 4) {
            const float4 v = *reinterpret_cast<const float4 *>(row + i);
            w[i] = v.x; w[i + 1] = v.y; w[i + 2] = v.z; w[i + 3] = v.w;
        }
    }
    __syncthreads();
    int T[GRU_R], col[GRU_R], Tmax = 0;
#pragma unroll
    for (int r = 0; r < GRU_R; r++) { T[r] = s_T[r]; col[r] = s_col[r]; Tmax = max(Tmax, T[r]); }

    float xn[GRU_R];
#pragma unroll
    for (int r = 0; r < GRU_R; r++) {
        const int t = backward ? (T[r] - 1) : 0;
        xn[r] = (T[r] > 0) ? Xin[(size_t)(col[r] + t) * (3 * H) + k] : 0.0f;
    }

    for (int s = 0; s < Tmax; s++) {
        const int cur = s & 1;
        float x[GRU_R];
#pragma unroll
        for (int r = 0; r < GRU_R; r++) x[r] = xn[r];
        if (s + 1 < Tmax) {
#pragma unroll
            for (int r = 0; r < GRU_R; r++) {
                const int t = backward ? (T[r] - 2 - s) : (s + 1);
                xn[r] = (s + 1 < T[r]) ? Xin[(size_t)(col[r] + t) * (3 * H) + k] : 0.0f;
            }
        }
        float acc[GRU_R];
#pragma unroll
        for (int r = 0; r < GRU_R; r++) acc[r] = 0.0f;

        if (k < 2 * H) {
#pragma unroll
            for (int i = 0; i < H; i++) {
                const float4 ha = *reinterpret_cast<const float4 *>(&hs[cur][i][0]);
                const float4 hb = *reinterpret_cast<const float4 *>(&hs[cur][i][4]);
                acc[0] = fmaf(w[i], ha.x, acc[0]); acc[1] = fmaf(w[i], ha.y, acc[1]);
                acc[2] = fmaf(w[i], ha.z, acc[2]); acc[3] = fmaf(w[i], ha.w, acc[3]);
                acc[4] = fmaf(w[i], hb.x, acc[4]); acc[5] = fmaf(w[i], hb.y, acc[5]);
                acc[6] = fmaf(w[i], hb.z, acc[6]); acc[7] = fmaf(w[i], hb.w, acc[7]);
            }
            if (k < H) {
#pragma unroll
                for (int r = 0; r < GRU_R; r++) zs[k][r] = logistic_cephes(x[r] + acc[r]);
            } else {
#pragma unroll
                for (int r = 0; r < GRU_R; r++) rhs[k - H][r] = logistic_cephes(x[r] + acc[r]) * hs[cur][k - H][r];
            }
        }
        __syncthreads();
        if (k >= 2 * H) {
            const int j = k - 2 * H;
#pragma unroll
            for (int i = 0; i < H; i++) {
                const float4 ha = *reinterpret_cast<const float4 *>(&rhs[i][0]);
                const float4 hb = *reinterpret_cast<const float4 *>(&rhs[i][4]);
                acc[0] = fmaf(w[i], ha.x, acc[0]); acc[1] = fmaf(w[i], ha.y, acc[1]);
                acc[2] = fmaf(w[i], ha.z, acc[2]); acc[3] = fmaf(w[i], ha.w, acc[3]);
                acc[4] = fmaf(w[i], hb.x, acc[4]); acc[5] = fmaf(w[i], hb.y, acc[5]);
                acc[6] = fmaf(w[i], hb.z, acc[6]); acc[7] = fmaf(w[i], hb.w, acc[7]);
            }
#pragma unroll
            for (int r = 0; r < GRU_R; r++) {
                const float cand = tanh_cephes(x[r] + acc[r]);
                const float z = zs[j][r];
                const float hn = z * hs[cur][j][r] + (1.0f - z) * cand;
                hs[cur ^ 1][j][r] = hn;
                if (s < T[r]) {
                    const int t = backward ? (T[r] - 1 - s) : s;
                    const size_t o = (size_t)(col[r] + t) * H + j;
                    out[o] = (resid != nullptr) ? hn + resid[o] : hn;
                }
            }
        }
        __syncthreads();
    }
}

void launch_gru_scan_ffma(const float *Xin, const float *sW, const float *sW2, const float *resid,
                          float *out, const BatchDims &d, int H, int backward, cudaStream_t s) {
    const int grid = (d.nread + GRU_R - 1) / GRU_R;
    if (H == 96)
        gru_scan_ffma_kernel<96><<<grid, 288, 0, s>>>(Xin, sW, sW2, resid, out, d, backward);
    else if (H == 112)
        gru_scan_ffma_kernel<112><<<grid, 336, 0, s>>>(Xin, sW, sW2, resid, out, d, backward);
}

// ---------------------------------------------------------------------------------
// CRF: global normalisation and Viterbi, one warp per read
// ---------------------------------------------------------------------------------
// Lane l < 25 owns transition (to = l / 5, from = l % 5).  The 5-vector of forward
// scores is replicated in every lane.
__device__ __forceinline__ float pick5(const float (&v)[5], int i) {
    float r = v[0];
    r = (i == 1) ? v[1] : r; r = (i == 2) ? v[2] : r; r = (i == 3) ? v[3] : r; r = (i == 4) ? v[4] : r;
    return r;
}

// ---------------------------------------------------------------------------------
// CRF (rnnrf_r94), second generation: transitions staged through shared memory
// ---------------------------------------------------------------------------------
// One warp per read.  Lanes 0..4 each own a destination state and evaluate its five incoming
// transitions locally, so a step costs five shuffles (broadcast of the new forward vector) instead of
// ten plus a serial log-sum-exp chain; all 32 lanes stream the transition rows in tiles of CRF_TILE
// steps (coalesced float4 loads, one tile ahead), which hides the HBM latency that the per-step loads of
// the first generation exposed.  The backtrace walks traceback tiles staged the same way.
constexpr int CRF_TILE = 32;
constexpr int CRF_WARPS = 4;

// crf_partition_function + globalnorm (src/layers.c:835-889).
__global__ void __launch_bounds__(32 * CRF_WARPS)
globalnorm_v2_kernel(float *__restrict__ trans, BatchDims d, int ostride) {
    __shared__ __align__(16) float tile[CRF_WARPS][2][CRF_TILE * 28];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int r = blockIdx.x * CRF_WARPS + warp;
    if (r >= d.nread || ostride != 28) return;
    const int T = d.nblock[r];
    float *tr = trans + (size_t)d.col_off[r] * 28;
    const int to = (lane < 5) ? lane : 0;
    float prev[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    constexpr int F4 = CRF_TILE * 28 / 4 / 32;          // float4 per lane per tile = 7
    float4 stage[F4];
    auto load_tile = [&](int t0) {
        const float4 *src = reinterpret_cast<const float4 *>(tr + (size_t)t0 * 28);
        const int nvalid4 = min(CRF_TILE, T - t0) * 7;   // 28 floats = 7 float4 per step
#pragma unroll
        for (int i = 0; i < F4; i++) {
            const int idx = i * 32 + lane;
            stage[i] = (idx < nvalid4) ? src[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    if (T > 0) load_tile(0);
    for (int t0 = 0, k = 0; t0 < T; t0 += CRF_TILE, k ^= 1) {
        float4 *dst = reinterpret_cast<float4 *>(tile[warp][k]);
#pragma unroll
        for (int i = 0; i < F4; i++) dst[i * 32 + lane] = stage[i];
        __syncwarp();
        if (t0 + CRF_TILE < T) load_tile(t0 + CRF_TILE);
        const int nt = min(CRF_TILE, T - t0);
        for (int t = 0; t < nt; t++) {
            const float *e = &tile[warp][k][t * 28 + to * 5];
            // the reference's nested pairwise log-sum-exp, in its order: the forward variables grow to ~5 T, where an
            // fp32 ulp is 0.03 for T = 80 000 -- any other summation form drifts from the reference's logZ by a
            // random walk of those roundings (measured: 14 on the longest bundled read)
            float v = e[0] + prev[0];
#pragma unroll
            for (int f = 1; f < 5; f++) v = logsumexp2(v, e[f] + prev[f]);
#pragma unroll
            for (int q = 0; q < 5; q++) prev[q] = __shfl_sync(0xffffffffu, v, q);
        }
        __syncwarp();
    }
    float logZ = prev[0];
#pragma unroll
    for (int q = 1; q < 5; q++) logZ = logsumexp2(logZ, prev[q]);
    logZ = logZ / (float)T;
    // second pass: subtract from the 25 real rows of every column (padding lanes stay zero)
    const size_t n = (size_t)T * 28;
    for (size_t i = lane; i < n; i += 32)
        if ((i % 28) < 25) tr[i] -= logZ;
}

void launch_globalnorm(float *trans, const BatchDims &d, int ostride, cudaStream_t s) {
    globalnorm_v2_kernel<<<(d.nread + CRF_WARPS - 1) / CRF_WARPS, 32 * CRF_WARPS, 0, s>>>(trans, d, ostride);
}

// decode_crf (src/decode.c:836-893): the reference's arithmetic and tie-breaking.
__global__ void __launch_bounds__(32 * CRF_WARPS)
decode_crf_v2_kernel(const float *__restrict__ trans, BatchDims d, int ostride, uint8_t *__restrict__ tb,
                     int *__restrict__ path, float *__restrict__ score) {
    __shared__ __align__(16) float tile[CRF_WARPS][2][CRF_TILE * 28];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int r = blockIdx.x * CRF_WARPS + warp;
    if (r >= d.nread || ostride != 28) return;
    const int T = d.nblock[r];
    const float *tr = trans + (size_t)d.col_off[r] * 28;
    uint8_t *tbr = tb + (size_t)d.col_off[r] * 8;
    const int to = (lane < 5) ? lane : 0;
    float prev[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    constexpr int F4 = CRF_TILE * 28 / 4 / 32;
    float4 stage[F4];
    auto load_tile = [&](int t0) {
        const float4 *src = reinterpret_cast<const float4 *>(tr + (size_t)t0 * 28);
        const int nvalid4 = min(CRF_TILE, T - t0) * 7;
#pragma unroll
        for (int i = 0; i < F4; i++) {
            const int idx = i * 32 + lane;
            stage[i] = (idx < nvalid4) ? src[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    if (T > 0) load_tile(0);
    for (int t0 = 0, k = 0; t0 < T; t0 += CRF_TILE, k ^= 1) {
        float4 *dst = reinterpret_cast<float4 *>(tile[warp][k]);
#pragma unroll
        for (int i = 0; i < F4; i++) dst[i * 32 + lane] = stage[i];
        __syncwarp();
        if (t0 + CRF_TILE < T) load_tile(t0 + CRF_TILE);
        const int nt = min(CRF_TILE, T - t0);
        for (int t = 0; t < nt; t++) {
            const float *e = &tile[warp][k][t * 28 + to * 5];
            float best = e[0] + prev[0];
            int arg = 0;
#pragma unroll
            for (int f = 1; f < 5; f++) {
                const float c = e[f] + prev[f];
                if (c > best) { best = c; arg = f; }       // strict: lowest `from` wins ties
            }
            if (lane < 8) tbr[(size_t)(t0 + t) * 8 + lane] = (lane < 5) ? (uint8_t)arg : (uint8_t)0;   // 8-byte rows: pad written too
#pragma unroll
            for (int q = 0; q < 5; q++) prev[q] = __shfl_sync(0xffffffffu, best, q);
        }
        __syncwarp();
    }
    int last = 0;
    for (int q = 1; q < 5; q++) if (prev[q] > prev[last]) last = q;
    int *p = path + d.col_off[r] + r;
    if (lane == 0) { score[r] = pick5(prev, last); p[T] = last; }
    // backtrace: traceback rows staged CRF_BT steps at a time (8 bytes per step) in the tile buffer
    constexpr int CRF_BT = 2 * CRF_TILE * 28 * 4 / 8;   // steps that fit the warp's tile buffers
    uint8_t *bt = reinterpret_cast<uint8_t *>(tile[warp][0]);
    __syncwarp();
    for (int hi = T; hi > 0; hi -= CRF_BT) {
        const int lo = max(hi - CRF_BT, 0);
        const int nrow = hi - lo;
        const uint2 *src = reinterpret_cast<const uint2 *>(tbr + (size_t)lo * 8);
        uint2 *dst2 = reinterpret_cast<uint2 *>(bt);
        for (int i = lane; i < nrow; i += 32) dst2[i] = src[i];
        __syncwarp();
        if (lane == 0) {
            for (int t = hi; t > lo; t--) {
                last = bt[(size_t)(t - 1 - lo) * 8 + last];
                p[t - 1] = last;
            }
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        __syncwarp();
    }
}

void launch_decode_crf(const float *trans, const BatchDims &d, int ostride, uint8_t *tb, int *path,
                       float *score, cudaStream_t s) {
    decode_crf_v2_kernel<<<(d.nread + CRF_WARPS - 1) / CRF_WARPS, 32 * CRF_WARPS, 0, s>>>(trans, d, ostride, tb, path, score);
}

// CRF head: C[col][0:25] = b + W^T x (feedforward_linear inside globalnorm, src/layers.c:874-880), M = 25 rows.
// One warp per column tile; lane m < M accumulates its row over k in ascending order (the same fused multiply-add
// sequence as affine_kernel); lanes M..ostride-1 write the zero padding.
constexpr int HEADS_COLS = 8;                          // columns per warp per iteration

__global__ void __launch_bounds__(256)
small_head_kernel(const float *__restrict__ X, int ncol, int K, const float *__restrict__ W, const float *__restrict__ b,
                  int M, float *__restrict__ C, int ldc) {
    extern __shared__ __align__(16) float sh[];
    float *Ws = sh;                                    // [K][32]: Ws[k * 32 + m]
    float *xs = sh + (size_t)K * 32;                   // per warp [HEADS_COLS][K]
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nwarp = blockDim.x / 32;
    for (int i = threadIdx.x; i < K * 32; i += blockDim.x) {
        const int k = i / 32, m = i % 32;
        Ws[i] = (m < M) ? W[(size_t)m * K + k] : 0.0f;
    }
    __syncthreads();
    float *xw = xs + (size_t)warp * HEADS_COLS * K;
    const float bm = (lane < M) ? b[lane] : 0.0f;
    for (int c0 = (blockIdx.x * nwarp + warp) * HEADS_COLS; c0 < ncol; c0 += gridDim.x * nwarp * HEADS_COLS) {
        const int nc = min(HEADS_COLS, ncol - c0);
        const float *src = X + (size_t)c0 * K;
        for (int i = lane; i < nc * K; i += 32) xw[i] = src[i];
        __syncwarp();
        float acc[HEADS_COLS];
#pragma unroll
        for (int j = 0; j < HEADS_COLS; j++) acc[j] = 0.0f;
        for (int k = 0; k < K; k++) {
            const float w = Ws[k * 32 + lane];
#pragma unroll
            for (int j = 0; j < HEADS_COLS; j++) acc[j] = fmaf(w, xw[j * K + k], acc[j]);
        }
        if (lane < ldc) {
#pragma unroll
            for (int j = 0; j < HEADS_COLS; j++)
                if (j < nc) C[(size_t)(c0 + j) * ldc + lane] = (lane < M) ? (bm + acc[j]) : 0.0f;
        }
        __syncwarp();
    }
}

void launch_small_head(const float *X, int ncol, int K, const float *W, const float *b, int M, float *C, int ldc,
                       cudaStream_t s) {
    const size_t smem = ((size_t)K * 32 + (size_t)8 * HEADS_COLS * K) * sizeof(float);
    const int grid = 148 * 4;
    small_head_kernel<<<grid, 256, smem, s>>>(X, ncol, K, W, b, M, C, ldc);
}


// ---------------------------------------------------------------------------------
// transducer Viterbi: one CTA per read, NH / 4 threads, 4 consecutive states per thread
// ---------------------------------------------------------------------------------
// Traceback is one byte per (block, state): 0 stay, 1+r step from r*NH/4 + i/4,
// 5+r skip from r*NH/16 + i/16, 21+r slip from r*NH/64 + i/64, 85 from the start state.
// The end state's predecessor is a separate int per block.  All comparisons are the
// strict ones of the reference, evaluated in its order (stay, step, skip, slip, start),
// and every maximum keeps the lowest index, so the result is bit-identical.
constexpr float DEC_BIG = 1.e30f;
enum { TB_STAY = 0, TB_STEP = 1, TB_SKIP = 5, TB_SLIP = 21, TB_START = 85 };

// ---------------------------------------------------------------------------------
// transducer Viterbi, second generation (same results, ~40 % fewer instructions per block)
// ---------------------------------------------------------------------------------
// Same thread <-> state mapping and traceback format as decode_transducer_kernel.  Differences:
//  * step / skip maxima are hierarchical: thread t computes m4[t] = max_r prev[r*NH/4 + t] once and
//    publishes (value, argmax); the 16-way skip maximum is then the max over four m4 entries.
//    argmax ties resolve to the lowest predecessor index exactly as the reference's ascending
//    scans do (r16 = 4a + b: the smallest r16 among the maximal entries).
//  * warp maxima use redux.sync (CREDUX.MAX.F32 / REDUX.MIN on sm_100a) instead of shuffle trees.
//  * only warp 0 tracks the end state.
//  * the backtrace stages 16 KB of traceback rows per chunk in shared memory with coalesced loads
//    from all threads, instead of one dependent HBM load per block from a single thread.
__device__ __forceinline__ float redux_max_f32(float v) {
    float m;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
    return m;
}
__device__ __forceinline__ int redux_min_s32(int v) {
    int m;
    asm volatile("redux.sync.min.s32 %0, %1, 0xffffffff;" : "=r"(m) : "r"(v));
    return m;
}

template <int NH>
__global__ void __launch_bounds__(NH / 4, (NH == 1024) ? 6 : 1)
decode_transducer_v2_kernel(const float *__restrict__ post, BatchDims d, int ostride, float stay_pen,
                            float skip_pen, float local_pen, int allow_slip, uint8_t *__restrict__ tb,
                            int *__restrict__ tb_end, int *__restrict__ path, float *__restrict__ score) {
    constexpr int NT = NH / 4;
    constexpr int NW = NT / 32;
    constexpr int BT_ROWS = 16384 / NH;                 // traceback rows staged per backtrace chunk (16 KB)
    constexpr int BIG_IDX = 0x7fffffff;
    constexpr int SC_BYTES = 2 * NH * 4, BT_BYTES = BT_ROWS * NH;
    // the score exchange buffers and the backtrace staging area are never live at the same time
    __shared__ __align__(16) uint8_t sc_or_bt[SC_BYTES > BT_BYTES ? SC_BYTES : BT_BYTES];
    __shared__ __align__(16) float2 m4s[NT];            // (max over the 4 step predecessors, its r) per suffix
    __shared__ float w_val[NW];
    __shared__ int w_idx[NW];
    __shared__ int s_last;
    float (*sc)[NH] = reinterpret_cast<float (*)[NH]>(sc_or_bt);
    uint8_t *bt_rows = sc_or_bt;

    const int r = blockIdx.x;
    const int t = threadIdx.x;
    const int lane = t % 32, warp = t / 32;
    const int T = d.nblock[r];
    const float *lp = post + (size_t)d.col_off[r] * ostride;
    uint8_t *tbr = tb + (size_t)d.col_off[r] * NH;
    int *tbe = tb_end + d.col_off[r];
    const float slip_pen = (float)(2.0 * (double)skip_pen);

    float cur[4] = {-DEC_BIG, -DEC_BIG, -DEC_BIG, -DEC_BIG};
    float curS = 0.0f, curE = -DEC_BIG;         // curS replicated in every thread, curE tracked by warp 0
    float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
    float nxt_stay = 0.0f;
    if (T > 0) {
        nxt = *reinterpret_cast<const float4 *>(lp + 4 * t);
        nxt_stay = lp[NH];
    }
    // running pointers: one add per block instead of re-deriving every address
    const float *lp_next = lp + ostride + 4 * t;        // this thread's 4 log-posteriors of block blk + 1
    const float *stay_next = lp + ostride + NH;
    uint8_t *tb_cur = tbr + 4 * t;
    int *tbe_cur = tbe;
    const float stay_c = -stay_pen, nlocal = -local_pen;
    for (int blk = 0; blk < T; blk++, lp_next += ostride, stay_next += ostride, tb_cur += NH, tbe_cur++) {
        const int buf = blk & 1;
        const float4 l4 = nxt;
        const float lstay = nxt_stay;
        if (blk + 1 < T) {
            nxt = *reinterpret_cast<const float4 *>(lp_next);
            nxt_stay = *stay_next;
        }
        // ---- phase A: publish the previous scores; this warp's best "enter end" candidate
        *reinterpret_cast<float4 *>(&sc[buf][4 * t]) = make_float4(cur[0], cur[1], cur[2], cur[3]);
        {
            const float v0 = cur[0] - local_pen, v1 = cur[1] - local_pen, v2 = cur[2] - local_pen, v3 = cur[3] - local_pen;
            const float wv = redux_max_f32(fmaxf(fmaxf(v0, v1), fmaxf(v2, v3)));
            int cand = BIG_IDX;
            cand = (v3 == wv) ? 4 * t + 3 : cand;
            cand = (v2 == wv) ? 4 * t + 2 : cand;
            cand = (v1 == wv) ? 4 * t + 1 : cand;
            cand = (v0 == wv) ? 4 * t : cand;
            const int wi = redux_min_s32(cand);          // lowest state index among the maximal ones
            if (lane == 0) { w_val[warp] = wv; w_idx[warp] = wi; }
        }
        __syncthreads();

        // ---- phase B: step maxima (one per thread); warp 0 also advances the end state
        const float *prev = sc[buf];
        float b4 = prev[t];
        int r4 = 0;
#pragma unroll
        for (int q = 1; q < 4; q++) {
            const float v = prev[q * (NH / 4) + t];
            if (b4 < v) { b4 = v; r4 = q; }
        }
        m4s[t] = make_float2(b4, __int_as_float(r4));
        if (warp == 0) {
            const float v = (lane < NW) ? w_val[lane] : -INFINITY;
            const int i = (lane < NW) ? w_idx[lane] : BIG_IDX;
            const float bv = redux_max_f32(v);
            const int bi = redux_min_s32((v == bv) ? i : BIG_IDX);
            float e = curE + fmaxf(-local_pen, lstay - stay_pen);
            int from = NH + 1;
            if (bv > e) { e = bv; from = bi; }
            if (lane == 0) *tbe_cur = from;
            curE = e;
        }
        __syncthreads();

        // ---- phase C: skip maximum from four step maxima, then the state updates
        float b16;
        int r16;
        {
            const float2 e0 = m4s[t / 4], e1 = m4s[(NH / 16) + t / 4], e2 = m4s[2 * (NH / 16) + t / 4], e3 = m4s[3 * (NH / 16) + t / 4];
            b16 = fmaxf(fmaxf(e0.x, e1.x), fmaxf(e2.x, e3.x));
            // predecessor block index r16 = 4 * a + b (a = e_b's own argmax): lowest among the maximal entries
            const int c0 = (e0.x == b16) ? 4 * __float_as_int(e0.y) : 99;
            const int c1 = (e1.x == b16) ? 4 * __float_as_int(e1.y) + 1 : 99;
            const int c2 = (e2.x == b16) ? 4 * __float_as_int(e2.y) + 2 : 99;
            const int c3 = (e3.x == b16) ? 4 * __float_as_int(e3.y) + 3 : 99;
            r16 = min(min(c0, c1), min(c2, c3));
        }
        float b64 = 0.0f;
        int r64 = 0;
        if (allow_slip) {
            b64 = prev[t / 16];
            for (int q = 1; q < 64; q++) {
                const float v = prev[q * (NH / 64) + t / 16];
                if (b64 < v) { b64 = v; r64 = q; }
            }
        }
        const float stay = lstay - stay_pen;
        const float lpj[4] = {l4.x, l4.y, l4.z, l4.w};
        uint32_t codes = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float s = cur[j] + stay;
            uint32_t code = TB_STAY;
            const float st = lpj[j] + b4;
            if (s < st) { s = st; code = TB_STEP + r4; }
            const float sk = (lpj[j] + b16) - skip_pen;
            if (s < sk) { s = sk; code = TB_SKIP + r16; }
            if (allow_slip) {
                const float sl = (lpj[j] + b64) - slip_pen;
                if (s < sl) { s = sl; code = TB_SLIP + r64; }
            }
            const float ss = curS + lpj[j];
            if (ss > s) { s = ss; code = TB_START; }
            cur[j] = s;
            codes |= code << (8 * j);
        }
        *reinterpret_cast<uint32_t *>(tb_cur) = codes;
        curS = curS + fmaxf(-local_pen, lstay - stay_pen);
    }

    // ---- final argmax over (states..., start, end): first maximum wins
    {
        const float wv = redux_max_f32(fmaxf(fmaxf(cur[0], cur[1]), fmaxf(cur[2], cur[3])));
        int cand = BIG_IDX;
        cand = (cur[3] == wv) ? 4 * t + 3 : cand;
        cand = (cur[2] == wv) ? 4 * t + 2 : cand;
        cand = (cur[1] == wv) ? 4 * t + 1 : cand;
        cand = (cur[0] == wv) ? 4 * t : cand;
        const int wi = redux_min_s32(cand);
        __syncthreads();
        if (lane == 0) { w_val[warp] = wv; w_idx[warp] = wi; }
    }
    __syncthreads();
    if (warp == 0) {
        const float v = (lane < NW) ? w_val[lane] : -INFINITY;
        const int i = (lane < NW) ? w_idx[lane] : BIG_IDX;
        float bv = redux_max_f32(v);
        int last = redux_min_s32((v == bv) ? i : BIG_IDX);
        if (curS > bv) { bv = curS; last = NH; }
        if (curE > bv) { bv = curE; last = NH + 1; }
        if (lane == 0) { score[r] = bv; s_last = last; }
    }
    __syncthreads();

    // ---- backtrace: chunks of BT_ROWS traceback rows staged in shared memory
    int *seq = path + d.col_off[r] + r;
    int last = s_last;
    for (int hi_blk = T; hi_blk > 0; hi_blk -= BT_ROWS) {
        const int lo_blk = max(hi_blk - BT_ROWS, 0);
        const int nrow = hi_blk - lo_blk;
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(tbr + (size_t)lo_blk * NH);
            uint4 *dst = reinterpret_cast<uint4 *>(bt_rows);
            for (int i = t; i < nrow * (NH / 16); i += NT) dst[i] = src[i];
        }
        __syncthreads();
        if (t == 0) {
            for (int blk = hi_blk - 1; blk >= lo_blk; blk--) {
                int out = -1;
                if (last == NH) {
                    out = NH;                                   // start stays in start
                } else if (last == NH + 1) {
                    out = NH + 1;
                    last = tbe[blk];
                } else {
                    const int code = bt_rows[(blk - lo_blk) * NH + last];
                    if (code != TB_STAY) {
                        out = last;
                        if (code >= TB_START) last = NH;
                        else if (code >= TB_SLIP) last = (code - TB_SLIP) * (NH / 64) + last / 64;
                        else if (code >= TB_SKIP) last = (code - TB_SKIP) * (NH / 16) + last / 16;
                        else last = (code - TB_STEP) * (NH / 4) + last / 4;
                    }
                }
                seq[blk + 1] = out;
            }
        }
        __syncthreads();
    }
    if (t == 0) {
        seq[0] = last;
        for (int i = 0; i < T; i++) { if (seq[i] == NH) seq[i] = -1; else break; }
        for (int i = T; i >= 0; i--) { if (seq[i] == NH + 1) seq[i] = -1; else break; }
    }
}

// SCRAPPIE_B200_DECODE (read once): "v2" forces the CTA-per-read kernel, "none" skips decoding (timing experiments,
// tools/exp_timeline.py).  Default: one warp per read with the scores in registers (kernels_decode.cu) for the
// 1024-history models without slip; the CTA-per-read kernel serves 4096 histories and slip.
static int decode_generation() {
    static const int gen = [] {
        const char *e = getenv("SCRAPPIE_B200_DECODE");
        return (e && 0 == strcmp(e, "none")) ? 0 : ((e && 0 == strcmp(e, "v2")) ? 2 : 4);
    }();
    return gen;
}

void launch_decode_transducer(const float *post, const BatchDims &d, int nstate, int ostride,
                              float stay_pen, float skip_pen, float local_pen, int allow_slip,
                              uint8_t *tb, int *tb_end, int *path, float *score, cudaStream_t s) {
    const int gen = decode_generation();
    const int nh = nstate - 1;
    if (gen == 0) return;
    if (gen == 4 && nh == 1024 && !allow_slip) {
        launch_decode_transducer_warp(post, d, ostride, stay_pen, skip_pen, local_pen, tb, tb_end, path, score, s);
        return;
    }
    if (nh == 1024)
        decode_transducer_v2_kernel<1024><<<d.nread, 256, 0, s>>>(post, d, ostride, stay_pen, skip_pen, local_pen,
                                                                 allow_slip, tb, tb_end, path, score);
    else if (nh == 4096)
        decode_transducer_v2_kernel<4096><<<d.nread, 1024, 0, s>>>(post, d, ostride, stay_pen, skip_pen, local_pen,
                                                                  allow_slip, tb, tb_end, path, score);
}

// ---------------------------------------------------------------------------------
// read finishing on the device: homopolymer fix-up + overlapper / crfpath_to_basecall
// ---------------------------------------------------------------------------------
// One warp per read walks that read's Viterbi path with the results of homopolymer_path
// (src/homopolymer.c:67-235: runs are detected on the ORIGINAL path, base-major, and applied in that
// order to the working copy) and overlapper (src/decode.c:367-382, :449-509) on the host.  Only the
// base strings travel back over PCIe.  The host implementations (host_decode.c) remain the library's
// single-read entry points and the cross-check of this kernel (tests/test_gpu_parity.py).
__device__ __forceinline__ int dev_kmer_shift(int prev, int next, int nkmer) {
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

__device__ __forceinline__ int dev_homopolymer_kmer(int base, int len) {
    int k = 0;
    for (int i = 0; i < len; i++) k = 4 * k + base;
    return k;
}

// expf as the host's libm evaluates it to within rounding: exp in double, rounded once to float
__device__ __forceinline__ double dev_expf_as_host(float x) { return (double)(float)exp((double)x); }

__device__ void dev_apply_run(const float *post, int ostride, int col0, int stay, int *pw, int start, int length, int state) {
    int nviterbi = 0;
    double expect = 0.0;
    for (int i = 0; i < length; i++) {
        const float *col = post + (size_t)(col0 + start + i - 1) * ostride;     // path[i] pairs with column i - 1
        const double ps = dev_expf_as_host(col[stay]);
        const double pr = dev_expf_as_host(col[state]);
        expect += pr / (pr + ps);
        if (pw[start + i] == state) nviterbi++;
    }
    const int nnew = (int)(expect + 0.5);
    if (nnew == nviterbi) return;
    for (int i = 0; i < length; i++) pw[start + i] = (i < nnew) ? state : -1;
}

// One WARP per read.  Run detection is evaluated for 32 positions at a time (ballot); hits are processed in ascending
// position order -- rule "XYYYY" before rule "ZXYYY" at the same position, bases outermost -- i.e. in exactly the
// host's order.  The overlapper is warp-parallel as well: 32 path entries per round, every move finds the k-mer before it
// with a ballot, the shifts are prefix-summed and every lane writes its own bases -- the same string as the serial
// loop of src/decode.c:449-509, 32 entries at a time.
// STAGED: original path, working copy and base string live in shared memory (reads of up to ~940 blocks: the fixed-length
// configurations).  Otherwise they stay in global memory (path_in, path_work, bases): no shared memory at all, so a
// batch of long reads neither waits for an SM with 200 KB free nor walks a 26 000-block path with a single thread.
constexpr int FIN_WARPS = 4;
constexpr size_t FIN_SMEM_MAX = 48 * 1024;

template <bool STAGED>
__global__ void __launch_bounds__(32 * FIN_WARPS)
finish_reads_warp_kernel(const float *__restrict__ post, BatchDims d, int nstate, int ostride, int head, int homopolymer,
                         int klen, int maxb, const int *__restrict__ path_in, int *__restrict__ path_work,
                         char *__restrict__ bases, int bases_stride, int *__restrict__ nbase_out) {
    extern __shared__ __align__(16) uint8_t fin_smem[];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int r = blockIdx.x * FIN_WARPS + warp;
    if (r >= d.nread) return;
    const int nb = d.nblock[r];
    const int col0 = d.col_off[r];
    const int *pin = path_in + col0 + r;
    char *out = bases + (size_t)r * bases_stride;
    const int *po;                                      // original path
    int *pw;                                            // working copy
    char *sb;                                           // base string under construction
    if (STAGED) {
        const int bcap = klen * (maxb + 1) + 1;
        const size_t per_warp = (size_t)2 * (maxb + 1) * sizeof(int) + (size_t)((bcap + 15) / 16 * 16);
        int *ps = reinterpret_cast<int *>(fin_smem + warp * per_warp);
        pw = ps + (maxb + 1);
        sb = reinterpret_cast<char *>(pw + (maxb + 1));
        for (int i = lane; i <= nb; i += 32) { const int v = pin[i]; ps[i] = v; pw[i] = v; }
        po = ps;
    } else {
        po = pin;
        pw = path_work + col0 + r;
        sb = out;
        if (head != 1)
            for (int i = lane; i <= nb; i += 32) pw[i] = pin[i];
    }
    __syncwarp();
    int nbase = 0;
    if (head == 1) {                                    // crfpath_to_basecall, src/decode.c:895-918
        for (int i0 = 0; i0 < nb; i0 += 32) {
            const int i = i0 + lane;
            const int st = (i < nb) ? po[i] : 4;
            const unsigned m = __ballot_sync(0xffffffffu, st < 4);
            if (st < 4) sb[nbase + __popc(m & ((1u << lane) - 1))] = "ACGT"[st];
            nbase += __popc(m);
        }
    } else {
        const int nkmer = nstate - 1;
        if (homopolymer == 1) {
            const int pathlen = nb;                     // homopolymer_path scans post->nc entries
            const int mod1 = 1 << (2 * (klen - 1)), mod2 = 1 << (2 * (klen - 2));
            const int stay = nstate - 1;
            for (int base = 0; base < 4; base++) {
                const int full = dev_homopolymer_kmer(base, klen);
                const int tail1 = dev_homopolymer_kmer(base, klen - 1);
                const int tail2 = dev_homopolymer_kmer(base, klen - 2);
                for (int i0 = 1; i0 < pathlen - 2; i0 += 32) {
                    const int i = i0 + lane;
                    bool c1 = false, c2 = false;
                    if (i < pathlen - 2) {
                        const int before = po[i - 1], here = po[i];
                        const bool ok = (before != -1) && ((here == -1) || (here == full));
                        c1 = ok && (before % mod1 == tail1) && (before != full);
                        c2 = ok && (before % mod2 == tail2) && (before % mod1 != tail1);
                    }
                    unsigned hits = __ballot_sync(0xffffffffu, c1 || c2);
                    const unsigned m1 = __ballot_sync(0xffffffffu, c1), m2 = __ballot_sync(0xffffffffu, c2);
                    while (hits) {                       // every lane walks the hits identically (warp-uniform)
                        const int l = __ffs(hits) - 1;
                        hits &= hits - 1;
                        const int ii = i0 + l;
                        if ((m1 >> l) & 1u) {
                            int e = ii + 1;
                            while (e < pathlen && (po[e] == -1 || po[e] == full)) e++;
                            if (lane == 0) dev_apply_run(post, ostride, col0, stay, pw, ii, e - ii, full);
                        }
                        if ((m2 >> l) & 1u) {
                            int j = ii;
                            while (j < pathlen && po[j] == -1) j++;
                            if (po[j] == full && j < pathlen - 1) {
                                int e = j + 1;
                                while (e < pathlen && (po[e] == -1 || po[e] == full)) e++;
                                if (lane == 0) dev_apply_run(post, ostride, col0, stay, pw, j, e - j, full);
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
        __syncwarp();
        // overlapper (src/decode.c:449-509), 32 path entries per round
        const int n = nb + 1;
        int first = n;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            const unsigned m = __ballot_sync(0xffffffffu, i < n && pw[i] >= 0);
            if (m) { first = i0 + __ffs(m) - 1; break; }
        }
        if (first == n) {
            nbase = -1;                                 // all stays: the host returns NULL
        } else {
            int prev_carry = pw[first];
            if (lane < klen) sb[klen - 1 - lane] = "ACGT"[(prev_carry >> (2 * lane)) & 3];
            int tail = klen - 1;
            for (int i0 = first + 1; i0 < n; i0 += 32) {
                const int i = i0 + lane;
                const int cur = (i < n) ? pw[i] : -1;
                const bool moved = cur >= 0;
                const unsigned m = __ballot_sync(0xffffffffu, moved);
                if (0 == m) continue;
                // the k-mer before this one: the nearest move below in this round, else the last move of earlier rounds
                const unsigned below = m & ((1u << lane) - 1u);
                const int src = below ? (31 - __clz(below)) : 0;
                const int pv = __shfl_sync(0xffffffffu, cur, src);
                const int prev = below ? pv : prev_carry;
                const int shift = moved ? dev_kmer_shift(prev, cur, nkmer) : 0;
                int incl = shift;
#pragma unroll
                for (int dl = 1; dl < 32; dl <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, dl);
                    if (lane >= dl) incl += t;
                }
                const int at = tail + (incl - shift);   // this move appends bases at at + 1 .. at + shift
                for (int j = 0, kmer = cur; j < shift; j++, kmer >>= 2) sb[at + shift - j] = "ACGT"[kmer & 3];
                tail += __shfl_sync(0xffffffffu, incl, 31);
                prev_carry = __shfl_sync(0xffffffffu, cur, 31 - __clz(m));
            }
            nbase = tail + 1;
        }
    }
    __syncwarp();
    const int nout = (nbase > 0) ? nbase : 0;
    if (STAGED)
        for (int i = lane; i < nout; i += 32) out[i] = sb[i];
    if (lane == 0) { out[nout] = 0; nbase_out[r] = nbase; }
}

void launch_finish_reads(const float *post, const BatchDims &d, int nstate, int ostride, int head, int homopolymer,
                         int klen, const int *path_in, int *path_work, char *bases, int bases_stride, int *nbase,
                         cudaStream_t s) {
    const int maxb = d.max_cols;
    const size_t per_warp = (size_t)2 * (maxb + 1) * sizeof(int) + (size_t)((klen * (maxb + 1) + 1 + 15) / 16 * 16);
    const size_t smem = per_warp * FIN_WARPS;
    const int grid = (d.nread + FIN_WARPS - 1) / FIN_WARPS;
    if (smem <= FIN_SMEM_MAX)       // within the default dynamic shared-memory limit: no function attribute needed
        finish_reads_warp_kernel<true><<<grid, 32 * FIN_WARPS, smem, s>>>(post, d, nstate, ostride, head, homopolymer, klen, maxb,
                                                                         path_in, path_work, bases, bases_stride, nbase);
    else
        finish_reads_warp_kernel<false><<<grid, 32 * FIN_WARPS, 0, s>>>(post, d, nstate, ostride, head, homopolymer, klen, maxb,
                                                                       path_in, path_work, bases, bases_stride, nbase);
}

// ---------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------
__global__ void gather_kernel(const float *__restrict__ post, int ostride, const int2 *__restrict__ idx, int n,
                              float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int2 cs = idx[i];                          // (column, state)
        out[i] = post[(size_t)cs.x * ostride + cs.y];
    }
}

void launch_gather(const float *post, int ostride, const int *col_state_pairs, int n, float *out, cudaStream_t s) {
    if (n > 0)
        gather_kernel<<<(n + 255) / 256, 256, 0, s>>>(post, ostride, reinterpret_cast<const int2 *>(col_state_pairs), n, out);
}

__global__ void flush_kernel(float *buf, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) buf[i] = 1.0f;
}

void launch_flush(float *buf, size_t nfloat, cudaStream_t s) { flush_kernel<<<148 * 8, 256, 0, s>>>(buf, nfloat); }

// Per-device function attributes of this file's kernels (called once per engine, after cudaSetDevice).
int configure_v1_kernels() {
    return cudaFuncSetAttribute(finish_reads_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FIN_SMEM_MAX) == cudaSuccess ? 0 : -1;
}

}  // namespace sb2
