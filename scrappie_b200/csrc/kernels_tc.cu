// Tensor-core (tcgen05 / TMEM) kernels for sm_100a.
//
//   gru_scan_tc_kernel   the recurrent part of a GRU layer (src/layers.c:373-527 in the
//                        reference): per time step two dependent products
//                        sW^T h (2H x N) and sW2^T (r*h) (H x N) as UMMA tiles with the
//                        gate math fused between them; weights stay resident in shared
//                        memory for the whole layer, accumulators live in TMEM.
//   tc_selftest_kernel   one UMMA tile product checked against the host (descriptor /
//                        layout validation and latency probe).
//
// Numerics: split-fp16 operands, three passes, fp32 accumulation (tc_common.cuh).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "device_math.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace sb2 {

using namespace tc;

// ---------------------------------------------------------------------------------
// self test: D[128][N] = A[128][K] * B[N][K]^T through the same operand path as the scan
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(160, 1)
tc_selftest_kernel(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ D, int K, int N,
                   int reps, long long *__restrict__ cycles) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t lboA = 128, sboA = (uint32_t)(K / 8) * 128;
    const uint32_t sboB = 128, lboB = 16u * N + 16u;
    const uint32_t tileA = 128u * K * 2u, tileB = (uint32_t)(K / 8) * lboB;
    uint8_t *a_hi = smem, *a_lo = smem + tileA, *b_hi = smem + 2 * tileA, *b_lo = b_hi + tileB;
    uint64_t *bars = reinterpret_cast<uint64_t *>(b_lo + tileB + 16);
    bars = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(bars) + 15) & ~(uintptr_t)15);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);

    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    for (int i = tid; i < 128 * K; i += blockDim.x) {
        const int m = i / K, k = i % K;
        __half hi, lo;
        split_fp16(A[i], hi, lo);
        const uint32_t off = canon_off(m, k, lboA, sboA);
        *reinterpret_cast<__half *>(a_hi + off) = hi;
        *reinterpret_cast<__half *>(a_lo + off) = lo;
    }
    for (int i = tid; i < N * K; i += blockDim.x) {
        const int n = i / K, k = i % K;
        __half hi, lo;
        split_fp16(B[i], hi, lo);
        const uint32_t off = canon_off(n, k, lboB, sboB);
        *reinterpret_cast<__half *>(b_hi + off) = hi;
        *reinterpret_cast<__half *>(b_lo + off) = lo;
    }
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    const uint32_t ncols = (N <= 32) ? 32 : 64;
    if (warp == 0) tmem_alloc(tmem_slot, ncols);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 4) {
        const uint32_t idesc = umma_idesc_f16(128, N);
        const uint64_t dA_hi = umma_desc(smem_u32(a_hi), lboA, sboA), dA_lo = umma_desc(smem_u32(a_lo), lboA, sboA);
        const uint64_t dB_hi = umma_desc(smem_u32(b_hi), lboB, sboB), dB_lo = umma_desc(smem_u32(b_lo), lboB, sboB);
        const int nk = K / 16;
        const uint64_t ka = (2 * lboA) >> 4, kb = (2 * lboB) >> 4;
        const long long t0 = clock64();
        long long t_issue = 0;
        for (int rep = 0; rep < reps; rep++) {
            const long long ta = clock64();
            if (elect_one()) {
                for (int ks = 0; ks < nk; ks++) umma_f16(tmem, dA_hi + ks * ka, dB_hi + ks * kb, idesc, ks > 0);
                for (int ks = 0; ks < nk; ks++) umma_f16(tmem, dA_lo + ks * ka, dB_hi + ks * kb, idesc, 1);
                for (int ks = 0; ks < nk; ks++) umma_f16(tmem, dA_hi + ks * ka, dB_lo + ks * kb, idesc, 1);
                umma_commit(&bars[0]);
            }
            __syncwarp();
            t_issue += clock64() - ta;
            mbar_wait(&bars[0], rep & 1);
        }
        const long long t1 = clock64();
        if (elect_one()) {
            cycles[0] = t1 - t0;
            cycles[1] = t_issue;
            umma_commit(&bars[1]);
        }
        __syncwarp();
    } else {
        mbar_wait(&bars[1], 0);
        tc_fence_after();
        const long long ta = clock64();
        for (int c0 = 0; c0 < N; c0 += 8) {
            float v[8];
            tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; j++) D[(size_t)tid * N + c0 + j] = v[j] * RESULT_SCALE;
        }
        if (tid == 0) cycles[2] = clock64() - ta;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, ncols);
}

int launch_tc_selftest(const float *A, const float *B, float *D, int K, int N, int reps, long long *cycles,
                       cudaStream_t s) {
    const size_t smem = 2 * (size_t)128 * K * 2 + 2 * (size_t)(K / 8) * (16 * N + 16) + 256;
    cudaError_t e = cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -1;
    tc_selftest_kernel<<<1, 160, smem, s>>>(A, B, D, K, N, reps, cycles);
    return 0;
}

// ---------------------------------------------------------------------------------
// GRU scan on tcgen05
// ---------------------------------------------------------------------------------
// One CTA = one tile of NR reads stepping together.  Warps 0..3 ("gate warps") own the
// accumulator rows (TMEM lane = hidden unit), warp 4 lane 0 issues the UMMAs.
//
// Shared memory: [weights image 6 tiles][B operands: h_hi h_lo rh_hi rh_lo][slack][barriers]
// The weights image is prepared once on the host (build_scan_image): per gate (z, r, c)
// a hi and a lo fp16 tile of H rows x H (K) in canonical layout, LBO 128, SBO (H/8)*128.
// UMMA uses M = 128, so rows H..127 of every tile read whatever follows it in shared
// memory; those accumulator lanes are never read back.
//
// Per step s (t = s forward, t = T-1-s backward):
//   UMMA  Dz, Dr  = Wz h, Wr h                  (3 passes x H/16 each)   -> commit g1
//   gates z = sig(xz + Dz), r = sig(xr + Dr); write (r*h) operand        -> arrive rh_ready
//   UMMA  Dc      = Wc (r*h)                                             -> commit g2
//   gates c = tanh(xc + Dc); h = z h + (1-z) c; store h; write h operand -> arrive h_ready
template <int H>
struct ScanLayout {
    static constexpr uint32_t LBO_A = 128;
    static constexpr uint32_t SBO_A = (H / 8) * 128;
    static constexpr uint32_t TILE_A = H * H * 2;
    static constexpr uint32_t WEIGHTS = 6 * TILE_A;
};

size_t scan_image_bytes(int H) { return (size_t)6 * H * H * 2; }

// Host: build the shared-memory image of a layer's recurrent weights.
// sW: [2H][H] (row = output unit: z rows then r rows), sW2: [H][H].
void build_scan_image(const float *sW, const float *sW2, int H, uint8_t *img) {
    const uint32_t lbo = 128, sbo = (uint32_t)(H / 8) * 128, tile = (uint32_t)H * H * 2;
    for (int g = 0; g < 3; g++) {
        uint8_t *hi_t = img + (size_t)(2 * g) * tile, *lo_t = hi_t + tile;
        for (int m = 0; m < H; m++) {
            const float *row = (g < 2) ? (sW + (size_t)(g * H + m) * H) : (sW2 + (size_t)m * H);
            for (int k = 0; k < H; k++) {
                const float xs = row[k] * OPERAND_SCALE;
                const __half hi = __float2half_rn(xs);
                const __half lo = __float2half_rn(xs - __half2float(hi));
                const uint32_t off = canon_off(m, k, lbo, sbo);
                *reinterpret_cast<__half *>(hi_t + off) = hi;
                *reinterpret_cast<__half *>(lo_t + off) = lo;
            }
        }
    }
}

template <int H, int NR, bool FAST>
__global__ void __launch_bounds__(160, 1)
gru_scan_tc_kernel(const float *__restrict__ Xin, const uint8_t *__restrict__ wimg, const float *__restrict__ resid,
                   float *__restrict__ out, BatchDims d, int backward) {
    using L = ScanLayout<H>;
    constexpr int NM = (NR < 16) ? 16 : NR;            // UMMA N (M = 128 requires N % 16 == 0)
    constexpr uint32_t LBO_B = 16u * NM + 16u, SBO_B = 128u;
    constexpr uint32_t TILE_B = (H / 8) * LBO_B;
    constexpr uint32_t SLACK = ((128 - H) / 8) * L::SBO_A;
    constexpr int NKS = H / 16;
    constexpr int NGW = (H + 31) / 32;                  // gate warps that own valid rows
    constexpr uint32_t TCOLS = (3 * NM <= 32) ? 32 : (3 * NM <= 64 ? 64 : (3 * NM <= 128 ? 128 : 256));

    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *w_img = smem;
    uint8_t *b_h_hi = smem + L::WEIGHTS, *b_h_lo = b_h_hi + TILE_B, *b_rh_hi = b_h_lo + TILE_B, *b_rh_lo = b_rh_hi + TILE_B;
    uint64_t *bars = reinterpret_cast<uint64_t *>(b_rh_lo + TILE_B + SLACK);    // 16-byte aligned by construction
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4);
    uint64_t *bar_g1 = &bars[0], *bar_g2 = &bars[1], *bar_rh = &bars[2], *bar_h = &bars[3];

    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int r0 = blockIdx.x * NR;

    // ---- one-time setup ---------------------------------------------------------
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(wimg);
        uint4 *dst = reinterpret_cast<uint4 *>(w_img);
        for (uint32_t i = tid; i < L::WEIGHTS / 16; i += blockDim.x) dst[i] = src[i];
        uint4 *zb = reinterpret_cast<uint4 *>(b_h_hi);
        for (uint32_t i = tid; i < (4 * TILE_B + SLACK) / 16; i += blockDim.x) zb[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        mbar_init(bar_g1, 1);
        mbar_init(bar_g2, 1);
        mbar_init(bar_rh, NGW);
        mbar_init(bar_h, NGW);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, TCOLS);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    int T[NR], col[NR], Tmax = 0;
#pragma unroll
    for (int n = 0; n < NR; n++) {
        const int r = r0 + n;
        T[n] = (r < d.nread) ? d.nblock[r] : 0;
        col[n] = (r < d.nread) ? d.col_off[r] : 0;
        Tmax = max(Tmax, T[n]);
    }

    if (warp == 4) {
        // ---- UMMA issuer: the whole warp runs the loop, one elected lane issues --------
        const uint32_t idesc = umma_idesc_f16(128, NM);
        const uint64_t dW = umma_desc(smem_u32(w_img), L::LBO_A, L::SBO_A);          // tile i at + i * TA
        const uint64_t dB = umma_desc(smem_u32(b_h_hi), LBO_B, SBO_B);               // h_hi, h_lo, rh_hi, rh_lo at + i * TB
        constexpr uint64_t TA = L::TILE_A >> 4, TB = TILE_B >> 4;
        constexpr uint64_t KA = (2 * L::LBO_A) >> 4, KB = (2 * LBO_B) >> 4;
        for (int s = 0; s < Tmax; s++) {
            if (s > 0) mbar_wait(bar_h, (s - 1) & 1);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int g = 0; g < 2; g++) {
                    const uint32_t dcol = tmem + g * NM;
                    const uint64_t w_hi = dW + (2 * g) * TA, w_lo = dW + (2 * g + 1) * TA;
#pragma unroll
                    for (int ks = 0; ks < NKS; ks++) umma_f16(dcol, w_hi + ks * KA, dB + ks * KB, idesc, ks > 0);
#pragma unroll
                    for (int ks = 0; ks < NKS; ks++) umma_f16(dcol, w_lo + ks * KA, dB + ks * KB, idesc, 1);
#pragma unroll
                    for (int ks = 0; ks < NKS; ks++) umma_f16(dcol, w_hi + ks * KA, dB + TB + ks * KB, idesc, 1);
                }
                umma_commit(bar_g1);
            }
            __syncwarp();
            mbar_wait(bar_rh, s & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t dcol = tmem + 2 * NM;
                const uint64_t w_hi = dW + 4 * TA, w_lo = dW + 5 * TA;
#pragma unroll
                for (int ks = 0; ks < NKS; ks++) umma_f16(dcol, w_hi + ks * KA, dB + 2 * TB + ks * KB, idesc, ks > 0);
#pragma unroll
                for (int ks = 0; ks < NKS; ks++) umma_f16(dcol, w_lo + ks * KA, dB + 2 * TB + ks * KB, idesc, 1);
#pragma unroll
                for (int ks = 0; ks < NKS; ks++) umma_f16(dcol, w_hi + ks * KA, dB + 3 * TB + ks * KB, idesc, 1);
                umma_commit(bar_g2);
            }
            __syncwarp();
        }
    } else if (warp < NGW) {
        // ---- gate warps -------------------------------------------------------------
        const int j = tid;                              // hidden unit = accumulator row = TMEM lane
        const bool valid = j < H;
        const int jj = valid ? j : 0;
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        float h[NR], xz[NR], xr[NR], xc[NR], rs[NR];
#pragma unroll
        for (int n = 0; n < NR; n++) { h[n] = 0.0f; rs[n] = 0.0f; }
        auto load_x = [&](int s) {
#pragma unroll
            for (int n = 0; n < NR; n++) {
                if (s < T[n]) {
                    const int t = backward ? (T[n] - 1 - s) : s;
                    const float *x = Xin + (size_t)(col[n] + t) * (3 * H) + jj;
                    xz[n] = x[0]; xr[n] = x[H]; xc[n] = x[2 * H];
                    if (resid != nullptr) rs[n] = resid[(size_t)(col[n] + t) * H + jj];
                } else {
                    xz[n] = 0.0f; xr[n] = 0.0f; xc[n] = 0.0f;
                }
            }
        };
        load_x(0);
        for (int s = 0; s < Tmax; s++) {
            float cz[NR], cr[NR], cc[NR], crs[NR];
#pragma unroll
            for (int n = 0; n < NR; n++) { cz[n] = xz[n]; cr[n] = xr[n]; cc[n] = xc[n]; crs[n] = rs[n]; }
            if (s + 1 < Tmax) load_x(s + 1);            // prefetch next step's inputs

            mbar_wait(bar_g1, s & 1);
            tc_fence_after();
            float gz[NR];
#pragma unroll
            for (int c0 = 0; c0 < NR; c0 += 8) {
                float vz[8], vr[8];
                tmem_ld8(lane_base + c0, vz);
                tmem_ld8(lane_base + NM + c0, vr);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int n = c0 + q;
                    const float az = cz[n] + vz[q] * RESULT_SCALE, ar = cr[n] + vr[q] * RESULT_SCALE;
                    gz[n] = FAST ? logistic_fast(az) : logistic_cephes(az);
                    const float gr = FAST ? logistic_fast(ar) : logistic_cephes(ar);
                    __half hi, lo;
                    split_fp16(gr * h[n], hi, lo);
                    if (valid) {
                        const uint32_t off = canon_off(n, j, LBO_B, SBO_B);
                        *reinterpret_cast<__half *>(b_rh_hi + off) = hi;
                        *reinterpret_cast<__half *>(b_rh_lo + off) = lo;
                    }
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_rh);

            mbar_wait(bar_g2, s & 1);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < NR; c0 += 8) {
                float vc[8];
                tmem_ld8(lane_base + 2 * NM + c0, vc);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int n = c0 + q;
                    const float ac = cc[n] + vc[q] * RESULT_SCALE;
                    const float cand = FAST ? tanh_fast(ac) : tanh_cephes(ac);
                    const float hn = gz[n] * h[n] + (1.0f - gz[n]) * cand;
                    h[n] = hn;
                    __half hi, lo;
                    split_fp16(hn, hi, lo);
                    if (valid) {
                        const uint32_t off = canon_off(n, j, LBO_B, SBO_B);
                        *reinterpret_cast<__half *>(b_h_hi + off) = hi;
                        *reinterpret_cast<__half *>(b_h_lo + off) = lo;
                        if (s < T[n]) {
                            const int t = backward ? (T[n] - 1 - s) : s;
                            out[(size_t)(col[n] + t) * H + j] = (resid != nullptr) ? hn + crs[n] : hn;
                        }
                    }
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_h);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TCOLS);
}


// ---------------------------------------------------------------------------------
// GRU scan, weights resident in TMEM (A operand from tensor memory)
// ---------------------------------------------------------------------------------
// Reading the 128 x 16 fp16 A tile from shared memory costs ~32 cycles per UMMA (4 KB at
// 128 B/clk), which dominated the shared-memory-A variant above (54 UMMAs per step).  Here
// the six weight tiles (z, r, c) x (hi, lo) are written once into TMEM with tcgen05.st
// (lane = hidden unit, one 32-bit column = two consecutive K elements; H/2 columns per
// tile) and every UMMA takes A from TMEM; shared memory only holds the small B operands
// (state h and r*h, split fp16).  Pass order per product: lo*hi, hi*lo, then hi*hi -- the
// tensor core accumulates with truncation, so the small cross terms go in while the
// accumulator is still small.
//
// MATH: 0 = cephes-identical gates, 1 = SFU ex2/rcp, 2 = polynomial exp2 + refined rcp.
template <int MATH>
__device__ __forceinline__ float gate_sigmoid(float x) {
    if (MATH == 0) return logistic_cephes(x);
    if (MATH == 1) return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x));
    // 2^t, t = -x log2(e) clamped; n = rint(t) by the magic-number trick; degree-6 minimax on [-1/2, 1/2]
    const float t = fmaxf(fminf(x * -1.4426950408889634f, 126.0f), -126.0f);
    const float r = t + 12582912.0f;
    const float f = t - (r - 12582912.0f);
    float p = 0.00015337577497120947f;
    p = fmaf(p, f, 0.0013399859890341759f);
    p = fmaf(p, f, 0.009618519805371761f);
    p = fmaf(p, f, 0.05550329014658928f);
    p = fmaf(p, f, 0.24022646248340607f);
    p = fmaf(p, f, 0.6931471824645996f);
    p = fmaf(p, f, 1.0f);
    const float e = __int_as_float(__float_as_int(p) + ((__float_as_int(r) - 0x4B400000) << 23));
    const float dd = 1.0f + e;
    const float q = rcp_approx(dd);
    return fmaf(q, fmaf(-dd, q, 1.0f), q);              // one Newton step
}
template <int MATH>
__device__ __forceinline__ float gate_tanh(float x) {
    if (MATH == 0) return tanh_cephes(x);
    const float y = gate_sigmoid<MATH>(x + x);
    return (y + y) - 1.0f;
}

__device__ __forceinline__ uint32_t pack_half2(__half lo16, __half hi16) {
    return (uint32_t)__half_as_ushort(lo16) | ((uint32_t)__half_as_ushort(hi16) << 16);
}

template <int H, int NR, int MATH>
__global__ void __launch_bounds__(160, 1)
gru_scan_tmem_kernel(const float *__restrict__ Xin, const float *__restrict__ sW, const float *__restrict__ sW2,
                     const float *__restrict__ resid, float *__restrict__ out, BatchDims d, int backward) {
    constexpr int NM = (NR < 16) ? 16 : NR;             // UMMA N
    constexpr uint32_t LBO_B = 16u * NM + 16u, SBO_B = 128u;
    constexpr uint32_t TILE_B = (H / 8) * LBO_B;
    constexpr int NKS = H / 16;
    constexpr int NGW = (H + 31) / 32;
    constexpr uint32_t KH = H / 2;                      // TMEM columns per weight tile
    constexpr uint32_t ACC0 = 6 * KH;                   // accumulators: z, r, c
    constexpr uint32_t TCOLS = 512;
    static_assert(ACC0 + 3 * NM <= TCOLS, "TMEM budget");

    __shared__ __align__(128) uint8_t b_ops[4 * TILE_B];        // h_hi, h_lo, rh_hi, rh_lo
    __shared__ __align__(8) uint64_t bars[4];
    __shared__ uint32_t tmem_slot;
    uint8_t *b_h_hi = b_ops, *b_h_lo = b_ops + TILE_B, *b_rh_hi = b_ops + 2 * TILE_B, *b_rh_lo = b_ops + 3 * TILE_B;
    uint64_t *bar_g1 = &bars[0], *bar_g2 = &bars[1], *bar_rh = &bars[2], *bar_h = &bars[3];

    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int r0 = blockIdx.x * NR;

    for (uint32_t i = tid; i < 4 * TILE_B / 16; i += blockDim.x) reinterpret_cast<uint4 *>(b_ops)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(bar_g1, 1);
        mbar_init(bar_g2, 1);
        mbar_init(bar_rh, NGW);
        mbar_init(bar_h, NGW);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, TCOLS);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    // ---- weights -> TMEM (once per layer) -------------------------------------------
    if (warp < 4) {
        const int m = tid;
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int g = 0; g < 3; g++) {
            const float *row = (g < 2) ? (sW + (size_t)(g * H + (m < H ? m : 0)) * H) : (sW2 + (size_t)(m < H ? m : 0) * H);
#pragma unroll 1
            for (int kc = 0; kc < NKS; kc++) {
                uint32_t whi[8], wlo[8];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    float4 v = *reinterpret_cast<const float4 *>(row + kc * 16 + q * 4);
                    if (m >= H) v = make_float4(0.f, 0.f, 0.f, 0.f);
                    __half h0, l0, h1, l1, h2, l2, h3, l3;
                    split_fp16(v.x, h0, l0); split_fp16(v.y, h1, l1); split_fp16(v.z, h2, l2); split_fp16(v.w, h3, l3);
                    whi[2 * q] = pack_half2(h0, h1); whi[2 * q + 1] = pack_half2(h2, h3);
                    wlo[2 * q] = pack_half2(l0, l1); wlo[2 * q + 1] = pack_half2(l2, l3);
                }
                tmem_st8(lane_base + (2 * g) * KH + kc * 8, whi);
                tmem_st8(lane_base + (2 * g + 1) * KH + kc * 8, wlo);
            }
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    int T[NR], col[NR], Tmax = 0;
#pragma unroll
    for (int n = 0; n < NR; n++) {
        const int r = r0 + n;
        T[n] = (r < d.nread) ? d.nblock[r] : 0;
        col[n] = (r < d.nread) ? d.col_off[r] : 0;
        Tmax = max(Tmax, T[n]);
    }

    if (warp == 4) {
        // ---- UMMA issuer ----------------------------------------------------------------
        const uint32_t idesc = umma_idesc_f16(128, NM);
        const uint64_t dB = umma_desc(smem_u32(b_ops), LBO_B, SBO_B);       // h_hi, h_lo, rh_hi, rh_lo at + i * TB
        constexpr uint64_t TB = TILE_B >> 4, KB = (2 * LBO_B) >> 4;
        for (int s = 0; s < Tmax; s++) {
            if (s > 0) mbar_wait(bar_h, (s - 1) & 1);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int g = 0; g < 2; g++) {
                    const uint32_t dcol = tmem + ACC0 + g * NM;
                    const uint32_t w_hi = tmem + (2 * g) * KH, w_lo = tmem + (2 * g + 1) * KH;
#pragma unroll
                    for (int ks = 0; ks < NKS; ks++) umma_f16_ts(dcol, w_lo + ks * 8, dB + ks * KB, idesc, ks > 0);
#pragma unroll
                    for (int ks = 0; ks < NKS; ks++) umma_f16_ts(dcol, w_hi + ks * 8, dB + TB + ks * KB, idesc, 1);
#pragma unroll
                    for (int ks = 0; ks < NKS; ks++) umma_f16_ts(dcol, w_hi + ks * 8, dB + ks * KB, idesc, 1);
                }
                umma_commit(bar_g1);
            }
            __syncwarp();
            mbar_wait(bar_rh, s & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t dcol = tmem + ACC0 + 2 * NM;
                const uint32_t w_hi = tmem + 4 * KH, w_lo = tmem + 5 * KH;
#pragma unroll
                for (int ks = 0; ks < NKS; ks++) umma_f16_ts(dcol, w_lo + ks * 8, dB + 2 * TB + ks * KB, idesc, ks > 0);
#pragma unroll
                for (int ks = 0; ks < NKS; ks++) umma_f16_ts(dcol, w_hi + ks * 8, dB + 3 * TB + ks * KB, idesc, 1);
#pragma unroll
                for (int ks = 0; ks < NKS; ks++) umma_f16_ts(dcol, w_hi + ks * 8, dB + 2 * TB + ks * KB, idesc, 1);
                umma_commit(bar_g2);
            }
            __syncwarp();
        }
    } else if (warp < NGW) {
        // ---- gate warps -----------------------------------------------------------------
        const int j = tid;
        const bool valid = j < H;
        const int jj = valid ? j : 0;
        const uint32_t acc_base = tmem + ((uint32_t)(warp * 32) << 16) + ACC0;
        float h[NR], xz[NR], xr[NR], xc[NR], rs[NR];
#pragma unroll
        for (int n = 0; n < NR; n++) { h[n] = 0.0f; rs[n] = 0.0f; }
        auto load_x = [&](int s) {
#pragma unroll
            for (int n = 0; n < NR; n++) {
                if (s < T[n]) {
                    const int t = backward ? (T[n] - 1 - s) : s;
                    const float *x = Xin + (size_t)(col[n] + t) * (3 * H) + jj;
                    xz[n] = x[0]; xr[n] = x[H]; xc[n] = x[2 * H];
                    if (resid != nullptr) rs[n] = resid[(size_t)(col[n] + t) * H + jj];
                } else {
                    xz[n] = 0.0f; xr[n] = 0.0f; xc[n] = 0.0f;
                }
            }
        };
        load_x(0);
        for (int s = 0; s < Tmax; s++) {
            float cz[NR], cr[NR], cc[NR], crs[NR];
#pragma unroll
            for (int n = 0; n < NR; n++) { cz[n] = xz[n]; cr[n] = xr[n]; cc[n] = xc[n]; crs[n] = rs[n]; }
            if (s + 1 < Tmax) load_x(s + 1);

            mbar_wait(bar_g1, s & 1);
            tc_fence_after();
            float gz[NR];
#pragma unroll
            for (int c0 = 0; c0 < NR; c0 += 8) {
                float vz[8], vr[8];
                tmem_ld8(acc_base + c0, vz);
                tmem_ld8(acc_base + NM + c0, vr);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int n = c0 + q;
                    gz[n] = gate_sigmoid<MATH>(fmaf(vz[q], RESULT_SCALE, cz[n]));
                    const float gr = gate_sigmoid<MATH>(fmaf(vr[q], RESULT_SCALE, cr[n]));
                    __half hi, lo;
                    split_fp16(gr * h[n], hi, lo);
                    if (valid) {
                        const uint32_t off = canon_off(n, j, LBO_B, SBO_B);
                        *reinterpret_cast<__half *>(b_rh_hi + off) = hi;
                        *reinterpret_cast<__half *>(b_rh_lo + off) = lo;
                    }
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_rh);

            mbar_wait(bar_g2, s & 1);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < NR; c0 += 8) {
                float vc[8];
                tmem_ld8(acc_base + 2 * NM + c0, vc);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int n = c0 + q;
                    const float cand = gate_tanh<MATH>(fmaf(vc[q], RESULT_SCALE, cc[n]));
                    const float hn = gz[n] * h[n] + (1.0f - gz[n]) * cand;
                    h[n] = hn;
                    __half hi, lo;
                    split_fp16(hn, hi, lo);
                    if (valid) {
                        const uint32_t off = canon_off(n, j, LBO_B, SBO_B);
                        *reinterpret_cast<__half *>(b_h_hi + off) = hi;
                        *reinterpret_cast<__half *>(b_h_lo + off) = lo;
                        if (s < T[n]) {
                            const int t = backward ? (T[n] - 1 - s) : s;
                            out[(size_t)(col[n] + t) * H + j] = (resid != nullptr) ? hn + crs[n] : hn;
                        }
                    }
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_h);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

template <int H, int NR, int MATH>
static int launch_scan_tmem(const float *Xin, const float *sW, const float *sW2, const float *resid, float *out,
                            const BatchDims &d, int backward, cudaStream_t s) {
    // The kernel allocates all 512 TMEM columns, so two CTAs must never share an SM (the second
    // would block in tcgen05.alloc until the first retires).  Requesting more than half of
    // the SM's shared memory as (unused) dynamic shared memory guarantees one CTA per SM.
    constexpr int EXCLUSIVE_SMEM = 120 * 1024;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(gru_scan_tmem_kernel<H, NR, MATH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 EXCLUSIVE_SMEM) != cudaSuccess)
            return -1;
        configured = true;
    }
    const int grid = (d.nread + NR - 1) / NR;
    gru_scan_tmem_kernel<H, NR, MATH><<<grid, 160, EXCLUSIVE_SMEM, s>>>(Xin, sW, sW2, resid, out, d, backward);
    return 0;
}

int launch_gru_scan_tmem(const float *Xin, const float *sW, const float *sW2, const float *resid, float *out,
                         const BatchDims &d, int H, int backward, int math, cudaStream_t s) {
#define SB2_CASE(HH, MM) if (H == HH && math == MM) return launch_scan_tmem<HH, 8, MM>(Xin, sW, sW2, resid, out, d, backward, s)
    SB2_CASE(96, 0); SB2_CASE(96, 1); SB2_CASE(96, 2);
    SB2_CASE(112, 0); SB2_CASE(112, 1); SB2_CASE(112, 2);
#undef SB2_CASE
    return -1;
}


// ---------------------------------------------------------------------------------
// GRU scan v3: two-pass split arithmetic, reset gate first, eight gate warps
// ---------------------------------------------------------------------------------
// Same data flow as gru_scan_tmem_kernel (weights resident in TMEM, TS-mode UMMA) with the
// per-step dependency chain shortened:
//  * The B operand holds the hi AND the lo half of the state side by side in the UMMA N
//    dimension: rows 0..7 = fp16 hi of the 8 reads, rows 8..15 = fp16 lo.  One product is then
//    two passes (A = W_lo, A = W_hi) instead of three, D[:, n] + D[:, 8 + n] is the result, and
//    the N = 16 the instruction needs anyway is fully used.  (W_lo * h_lo comes for free.)
//  * The reset gate is issued and committed first; the update gate's UMMAs run on the tensor
//    pipe while the gate warps turn r into the (r * h) operand, and sigma(z) is evaluated while
//    the candidate's UMMAs run.
//  * Eight gate warps: TMEM lane quarter q = warp % 4 (hidden units 32q..32q+31), read group
//    cg = warp / 4 (reads 4cg..4cg+3), so a thread evaluates 4 reads instead of 8.
// At N = 16 a UMMA costs ~38 cycles whatever it computes (profiles/r5_summary.md), so the step
// time is (36 UMMAs) x 38 cycles plus the two hand-overs; see DESIGN.md.
template <int H, int MATH>
__global__ void __launch_bounds__(288, 1)
gru_scan_v3_kernel(const float *__restrict__ Xin, const float *__restrict__ sW, const float *__restrict__ sW2,
                   const float *__restrict__ resid, float *__restrict__ out, BatchDims d, int backward,
                   long long *__restrict__ trace, int dbg) {
    // diagnostic variants (timing only, wrong results): 1 no global IO, 2 no gate math, 4 no loads, 8 no stores
    const bool dbg_no_math = (trace != nullptr) && (dbg & 2);
    const bool dbg_no_ld = (trace != nullptr) && (dbg & 5);
    const bool dbg_no_st = (trace != nullptr) && (dbg & 9);
    // trace (diagnostic, normally null): CTA 0 records clock64() at the hand-over points of steps 100..103;
    // slots 0..3 issuer, 4..12 gate warp 0 (see tools/scan_trace.py)
#define SB2_TRACE(slot) do { if (trace != nullptr && blockIdx.x == 0 && lane == 0 && s >= 100 && s < 104) trace[(s - 100) * 16 + (slot)] = clock64(); } while (0)
    constexpr int NR = 8, NM = 16, RPT = 4;             // reads per CTA, UMMA N, reads per gate thread
    constexpr uint32_t LBO_B = 16u * NM + 16u, SBO_B = 128u;
    constexpr uint32_t TILE_B = (H / 8) * LBO_B;
    constexpr int NKS = H / 16;
    constexpr int NQ = (H + 31) / 32;                   // lane quarters that own hidden units
    constexpr int NGW = 2 * NQ;                         // active gate warps
    constexpr uint32_t KH = H / 2;                      // TMEM columns per weight tile
    constexpr uint32_t ACC0 = 6 * KH;                   // accumulators: r, z, c (16 columns each)
    constexpr uint32_t TCOLS = 512;
    static_assert(ACC0 + 3 * NM <= TCOLS, "TMEM budget");

    __shared__ __align__(128) uint8_t b_ops[2 * TILE_B];        // h [hi|lo], r*h [hi|lo]
    __shared__ __align__(8) uint64_t bars[5];
    __shared__ uint32_t tmem_slot;
    uint8_t *b_h = b_ops, *b_rh = b_ops + TILE_B;
    uint64_t *bar_r = &bars[0], *bar_z = &bars[1], *bar_c = &bars[2], *bar_rh = &bars[3], *bar_h = &bars[4];

    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int r0 = blockIdx.x * NR;

    for (uint32_t i = tid; i < 2 * TILE_B / 16; i += blockDim.x) reinterpret_cast<uint4 *>(b_ops)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(bar_r, 1);
        mbar_init(bar_z, 1);
        mbar_init(bar_c, 1);
        mbar_init(bar_rh, NGW);
        mbar_init(bar_h, NGW);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, TCOLS);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    // ---- weights -> TMEM (once per layer): tiles r_hi r_lo z_hi z_lo c_hi c_lo -------------
    if (warp < 4) {
        const int m = tid;
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int g = 0; g < 3; g++) {
            // tile order r, z, c; the reference stores z rows first, then r (src/layers.c:511-526)
            const int mm = (m < H) ? m : 0;
            const float *row = (g == 0) ? (sW + (size_t)(H + mm) * H) : ((g == 1) ? (sW + (size_t)mm * H) : (sW2 + (size_t)mm * H));
#pragma unroll 1
            for (int kc = 0; kc < NKS; kc++) {
                uint32_t whi[8], wlo[8];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    float4 v = *reinterpret_cast<const float4 *>(row + kc * 16 + q * 4);
                    if (m >= H) v = make_float4(0.f, 0.f, 0.f, 0.f);
                    __half h0, l0, h1, l1, h2, l2, h3, l3;
                    split_fp16(v.x, h0, l0); split_fp16(v.y, h1, l1); split_fp16(v.z, h2, l2); split_fp16(v.w, h3, l3);
                    whi[2 * q] = pack_half2(h0, h1); whi[2 * q + 1] = pack_half2(h2, h3);
                    wlo[2 * q] = pack_half2(l0, l1); wlo[2 * q + 1] = pack_half2(l2, l3);
                }
                tmem_st8(lane_base + (2 * g) * KH + kc * 8, whi);
                tmem_st8(lane_base + (2 * g + 1) * KH + kc * 8, wlo);
            }
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 8) {
        // ---- UMMA issuer --------------------------------------------------------------------
        int Tmax = 0;
        for (int n = 0; n < NR; n++) {
            const int r = r0 + n;
            if (r < d.nread) Tmax = max(Tmax, d.nblock[r]);
        }
        const uint32_t idesc = umma_idesc_f16(128, NM);
        const uint64_t dBh = umma_desc(smem_u32(b_h), LBO_B, SBO_B), dBrh = umma_desc(smem_u32(b_rh), LBO_B, SBO_B);
        constexpr uint64_t KB = (2 * LBO_B) >> 4;
        for (int s = 0; s < Tmax; s++) {
            if (s > 0) mbar_wait(bar_h, (s - 1) & 1);
            tc_fence_after();
            SB2_TRACE(0);
            if (elect_one()) {
#pragma unroll
                for (int g = 0; g < 2; g++) {
                    const uint32_t dcol = tmem + ACC0 + g * NM;
                    const uint32_t w_hi = tmem + (2 * g) * KH, w_lo = w_hi + KH;
#pragma unroll
                    for (int ks = 0; ks < NKS; ks++) umma_f16_ts(dcol, w_lo + ks * 8, dBh + ks * KB, idesc, ks > 0);
#pragma unroll
                    for (int ks = 0; ks < NKS; ks++) umma_f16_ts(dcol, w_hi + ks * 8, dBh + ks * KB, idesc, 1);
                    umma_commit(g == 0 ? bar_r : bar_z);
                }
            }
            __syncwarp();
            SB2_TRACE(1);
            mbar_wait(bar_rh, s & 1);
            tc_fence_after();
            SB2_TRACE(2);
            if (elect_one()) {
                const uint32_t dcol = tmem + ACC0 + 2 * NM;
                const uint32_t w_hi = tmem + 4 * KH, w_lo = w_hi + KH;
#pragma unroll
                for (int ks = 0; ks < NKS; ks++) umma_f16_ts(dcol, w_lo + ks * 8, dBrh + ks * KB, idesc, ks > 0);
#pragma unroll
                for (int ks = 0; ks < NKS; ks++) umma_f16_ts(dcol, w_hi + ks * 8, dBrh + ks * KB, idesc, 1);
                umma_commit(bar_c);
            }
            __syncwarp();
            SB2_TRACE(3);
        }
    } else if ((warp & 3) < NQ) {
        // ---- gate warps -----------------------------------------------------------------------
        const int q = warp & 3, cg = warp >> 2;
        const int j = q * 32 + lane;                    // hidden unit = accumulator row = TMEM lane
        const bool valid = j < H;
        const int jj = valid ? j : 0;
        const uint32_t acc_base = tmem + ((uint32_t)(q * 32) << 16) + ACC0 + cg * RPT;
        int T[RPT], col[RPT], Tmax = 0;
        for (int n = 0; n < NR; n++) {
            const int r = r0 + n;
            if (r < d.nread) Tmax = max(Tmax, d.nblock[r]);
        }
#pragma unroll
        for (int i = 0; i < RPT; i++) {
            const int r = r0 + cg * RPT + i;
            T[i] = (r < d.nread) ? d.nblock[r] : 0;
            col[i] = (r < d.nread) ? d.col_off[r] : 0;
        }
        // byte offsets of this thread's operand elements: row n = cg*4 + i (hi), + SBO_B (lo)
        const uint32_t op_off = (uint32_t)(j >> 3) * LBO_B + (uint32_t)(cg * RPT) * 16 + (uint32_t)(j & 7) * 2;
        float h[RPT], xz[RPT], xr[RPT], xc[RPT], rs[RPT];
#pragma unroll
        for (int i = 0; i < RPT; i++) { h[i] = 0.0f; rs[i] = 0.0f; }
        auto load_x = [&](int s) {
#pragma unroll
            for (int i = 0; i < RPT; i++) {
                if (s < T[i]) {
                    const int t = backward ? (T[i] - 1 - s) : s;
                    const float *x = Xin + (size_t)(col[i] + t) * (3 * H) + jj;
                    xz[i] = x[0]; xr[i] = x[H]; xc[i] = x[2 * H];
                    if (resid != nullptr) rs[i] = resid[(size_t)(col[i] + t) * H + jj];
                } else {
                    xz[i] = 0.0f; xr[i] = 0.0f; xc[i] = 0.0f;
                }
            }
        };
        load_x(0);
        for (int s = 0; s < Tmax; s++) {
            float cz[RPT], cr[RPT], cc[RPT], crs[RPT];
#pragma unroll
            for (int i = 0; i < RPT; i++) { cz[i] = xz[i]; cr[i] = xr[i]; cc[i] = xc[i]; crs[i] = rs[i]; }
            if (s + 1 < Tmax && !dbg_no_ld) load_x(s + 1);            // prefetch the next step's inputs

            // reset gate -> (r * h) operand
            mbar_wait(bar_r, s & 1);
            tc_fence_after();
            if (warp == 0) SB2_TRACE(4);
            {
                float a[4], b[4];
                tmem_ld4(acc_base, a);
                tmem_ld4(acc_base + NR, b);
                tmem_ld_wait();
                if (warp == 0) SB2_TRACE(5);
#pragma unroll
                for (int i = 0; i < RPT; i++) {
                    const float gr = dbg_no_math ? (a[i] + b[i]) : gate_sigmoid<MATH>(fmaf(a[i] + b[i], RESULT_SCALE, cr[i]));
                    __half hi, lo;
                    split_fp16(gr * h[i], hi, lo);
                    if (valid) {
                        *reinterpret_cast<__half *>(b_rh + op_off + i * 16) = hi;
                        *reinterpret_cast<__half *>(b_rh + op_off + i * 16 + SBO_B) = lo;
                    }
                }
            }
            if (warp == 0) SB2_TRACE(6);
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_rh);
            if (warp == 0) SB2_TRACE(7);

            // update gate (its UMMAs ran while the reset gate was being evaluated)
            float gz[RPT];
            mbar_wait(bar_z, s & 1);
            tc_fence_after();
            if (warp == 0) SB2_TRACE(8);
            {
                float a[4], b[4];
                tmem_ld4(acc_base + NM, a);
                tmem_ld4(acc_base + NM + NR, b);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < RPT; i++) gz[i] = dbg_no_math ? a[i] : gate_sigmoid<MATH>(fmaf(a[i] + b[i], RESULT_SCALE, cz[i]));
            }

            // candidate, state update, next step's operand
            if (warp == 0) SB2_TRACE(9);
            mbar_wait(bar_c, s & 1);
            tc_fence_after();
            if (warp == 0) SB2_TRACE(10);
            float hn[RPT];
            {
                float a[4], b[4];
                tmem_ld4(acc_base + 2 * NM, a);
                tmem_ld4(acc_base + 2 * NM + NR, b);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < RPT; i++) {
                    const float cand = dbg_no_math ? b[i] : gate_tanh<MATH>(fmaf(a[i] + b[i], RESULT_SCALE, cc[i]));
                    hn[i] = gz[i] * h[i] + (1.0f - gz[i]) * cand;
                    h[i] = hn[i];
                    __half hi, lo;
                    split_fp16(hn[i], hi, lo);
                    if (valid) {
                        *reinterpret_cast<__half *>(b_h + op_off + i * 16) = hi;
                        *reinterpret_cast<__half *>(b_h + op_off + i * 16 + SBO_B) = lo;
                    }
                }
            }
            if (warp == 0) SB2_TRACE(11);
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_h);
            if (warp == 0) SB2_TRACE(12);
            // results to HBM, off the critical path
            if (valid && !dbg_no_st) {
#pragma unroll
                for (int i = 0; i < RPT; i++) {
                    if (s < T[i]) {
                        const int t = backward ? (T[i] - 1 - s) : s;
                        out[(size_t)(col[i] + t) * H + j] = (resid != nullptr) ? hn[i] + crs[i] : hn[i];
                    }
                }
            }
        }
    }
#undef SB2_TRACE
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

template <int H, int MATH>
static int launch_scan_v3(const float *Xin, const float *sW, const float *sW2, const float *resid, float *out,
                          const BatchDims &d, int backward, long long *trace, int dbg, cudaStream_t s) {
    // all 512 TMEM columns are allocated: keep a second scan CTA off the SM (see launch_scan_tmem)
    constexpr int EXCLUSIVE_SMEM = 120 * 1024;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(gru_scan_v3_kernel<H, MATH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 EXCLUSIVE_SMEM) != cudaSuccess)
            return -1;
        configured = true;
    }
    const int grid = (d.nread + 7) / 8;
    gru_scan_v3_kernel<H, MATH><<<grid, 288, EXCLUSIVE_SMEM, s>>>(Xin, sW, sW2, resid, out, d, backward, trace, dbg);
    return 0;
}

int launch_gru_scan_v3(const float *Xin, const float *sW, const float *sW2, const float *resid, float *out,
                       const BatchDims &d, int H, int backward, int math, long long *trace, cudaStream_t s) {
    static int dbg = -1;
    if (dbg < 0) { const char *e = getenv("SCRAPPIE_B200_SCAN_DBG"); dbg = e ? atoi(e) : 0; }
#define SB2_CASE(HH, MM) if (H == HH && math == MM) return launch_scan_v3<HH, MM>(Xin, sW, sW2, resid, out, d, backward, trace, dbg, s)
    SB2_CASE(96, 0); SB2_CASE(96, 1); SB2_CASE(96, 2);
    SB2_CASE(112, 0); SB2_CASE(112, 1); SB2_CASE(112, 2);
#undef SB2_CASE
    return -1;
}


// ---------------------------------------------------------------------------------
// GRU scan v4: two independent read groups per CTA, gate math written for ILP
// ---------------------------------------------------------------------------------
// Measured on v3 (tools/scan_trace.py, profiles/): the UMMAs of a step are cheap (~10-14 cycles
// each to issue, ~300 cycles from first issue to the commit being seen); the step time is set by
// the CUDA-core side -- the gate polynomials are a long dependent chain, and a TMEM lane quarter
// can only be read by the warps of ONE scheduler (warp % 4), so all the math of 32 hidden units
// lands on one SMSP.  v4 therefore
//  * splits the 8 reads of a CTA into two groups of 4 with their own operands, accumulators,
//    barriers and issuer warp, so one group's gate math fills the other group's UMMA / hand-over
//    waits on the same schedulers;
//  * evaluates the 4 reads of a thread stage by stage (sigmoid4 / tanh4) so the four dependent
//    chains interleave;
//  * keeps hi and lo of a read in the same 8-row core-matrix group (rows i and 4 + i), so one
//    tcgen05.ld.x8 fetches both partial sums.
template <int MATH>
__device__ __forceinline__ void sigmoid4(const float (&x)[4], float (&y)[4]) {
    if (MATH == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) y[i] = logistic_cephes(x[i]);
    } else if (MATH == 1) {
        float e[4];
#pragma unroll
        for (int i = 0; i < 4; i++) e[i] = ex2_approx(-1.4426950408889634f * x[i]);
#pragma unroll
        for (int i = 0; i < 4; i++) y[i] = rcp_approx(1.0f + e[i]);
    } else if (MATH == 5) {
        // ex2.approx on the plainly rounded argument, Newton-refined reciprocal
        float e[4], d[4], q[4];
#pragma unroll
        for (int i = 0; i < 4; i++) e[i] = ex2_approx(fminf(x[i] * -1.4426950408889634f, 126.0f));
#pragma unroll
        for (int i = 0; i < 4; i++) d[i] = 1.0f + e[i];
#pragma unroll
        for (int i = 0; i < 4; i++) q[i] = rcp_approx(d[i]);
#pragma unroll
        for (int i = 0; i < 4; i++) y[i] = fmaf(q[i], fmaf(-d[i], q[i], 1.0f), q[i]);
    } else if (MATH == 3 || MATH == 4) {
        // SFU exponential with a compensated argument: t = -x log2(e) is formed as th + tl (tl = the rounding
        // error of the product plus the low part of the constant), 2^t = ex2(th) * (1 + ln2 * tl).  Removes the
        // |t| * 2^-24 argument error that dominates ex2.approx(x * log2e) for |x| > 2; ~3 ulp overall.
        float th[4], tl[4], e[4], d[4], q[4];
#pragma unroll
        for (int i = 0; i < 4; i++) th[i] = x[i] * -1.4426950216293335f;
#pragma unroll
        for (int i = 0; i < 4; i++) tl[i] = fmaf(x[i], -1.4426950216293335f, -th[i]);
#pragma unroll
        for (int i = 0; i < 4; i++) tl[i] = fmaf(x[i], -1.9259629911783985e-08f, tl[i]);
#pragma unroll
        for (int i = 0; i < 4; i++) e[i] = ex2_approx(fminf(th[i], 126.0f));       // keep 1 + e finite
#pragma unroll
        for (int i = 0; i < 4; i++) tl[i] = tl[i] * 0.6931471805599453f;
#pragma unroll
        for (int i = 0; i < 4; i++) d[i] = 1.0f + fmaf(e[i], tl[i], e[i]);
#pragma unroll
        for (int i = 0; i < 4; i++) q[i] = rcp_approx(d[i]);
#pragma unroll
        for (int i = 0; i < 4; i++) y[i] = (MATH == 4) ? q[i] : fmaf(q[i], fmaf(-d[i], q[i], 1.0f), q[i]);   // 4: no Newton step
    } else {
        float t[4], r[4], f[4], p[4], dd[4], q[4];
#pragma unroll
        for (int i = 0; i < 4; i++) t[i] = fmaxf(fminf(x[i] * -1.4426950408889634f, 126.0f), -126.0f);
#pragma unroll
        for (int i = 0; i < 4; i++) r[i] = t[i] + 12582912.0f;
#pragma unroll
        for (int i = 0; i < 4; i++) f[i] = t[i] - (r[i] - 12582912.0f);
#pragma unroll
        for (int i = 0; i < 4; i++) p[i] = fmaf(0.00015337577497120947f, f[i], 0.0013399859890341759f);
#pragma unroll
        for (int i = 0; i < 4; i++) p[i] = fmaf(p[i], f[i], 0.009618519805371761f);
#pragma unroll
        for (int i = 0; i < 4; i++) p[i] = fmaf(p[i], f[i], 0.05550329014658928f);
#pragma unroll
        for (int i = 0; i < 4; i++) p[i] = fmaf(p[i], f[i], 0.24022646248340607f);
#pragma unroll
        for (int i = 0; i < 4; i++) p[i] = fmaf(p[i], f[i], 0.6931471824645996f);
#pragma unroll
        for (int i = 0; i < 4; i++) p[i] = fmaf(p[i], f[i], 1.0f);
#pragma unroll
        for (int i = 0; i < 4; i++)
            dd[i] = 1.0f + __int_as_float(__float_as_int(p[i]) + ((__float_as_int(r[i]) - 0x4B400000) << 23));
#pragma unroll
        for (int i = 0; i < 4; i++) q[i] = rcp_approx(dd[i]);
#pragma unroll
        for (int i = 0; i < 4; i++) y[i] = fmaf(q[i], fmaf(-dd[i], q[i], 1.0f), q[i]);
    }
}
template <int MATH>
__device__ __forceinline__ void tanh4(const float (&x)[4], float (&y)[4]) {
    if (MATH == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) y[i] = tanh_cephes(x[i]);
    } else {
        float x2[4], s[4];
#pragma unroll
        for (int i = 0; i < 4; i++) x2[i] = x[i] + x[i];
        sigmoid4<MATH>(x2, s);
#pragma unroll
        for (int i = 0; i < 4; i++) y[i] = (s[i] + s[i]) - 1.0f;
    }
}

template <int H, int MATH, int NG>
__global__ void __launch_bounds__((H > 96) ? 512 : 128 * NG, 1)
gru_scan_v4_kernel(const float *__restrict__ Xin, const float *__restrict__ sW, const float *__restrict__ sW2,
                   const float *__restrict__ resid, float *__restrict__ out, BatchDims d, int backward,
                   long long *__restrict__ trace) {
#define SB2_TRACE(slot) do { if (trace != nullptr && blockIdx.x == 0 && grp == 0 && lane == 0 && s >= 100 && s < 104) trace[(s - 100) * 16 + (slot)] = clock64(); } while (0)
    constexpr int RPG = 4, NM = 16;                     // NG groups per CTA; reads per group, UMMA N
    static_assert(NG == 2 || (NG == 4 && H <= 96) || (NG == 3 && H > 96), "groups per CTA: TMEM columns and warp slots");
    constexpr uint32_t LBO_B = 16u * NM + 16u, SBO_B = 128u;
    constexpr uint32_t TILE_B = (H / 8) * LBO_B;
    constexpr int NKS = H / 16;
    constexpr int NQ = (H + 31) / 32;                   // lane quarters that own hidden units
    constexpr uint32_t KH = H / 2;                      // TMEM columns per weight tile
    constexpr uint32_t ACC0 = 6 * KH;                   // accumulators: per group r, z, c (16 columns each)
    constexpr uint32_t TCOLS = 512;
    static_assert(ACC0 + NG * 3 * NM <= TCOLS, "TMEM budget");

    __shared__ __align__(128) uint8_t b_ops[NG * 2 * TILE_B];   // per group: h [hi|lo], r*h [hi|lo]
    __shared__ __align__(8) uint64_t bars[NG * 5];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int r0 = blockIdx.x * (NG * RPG);

    for (uint32_t i = tid; i < NG * 2 * TILE_B / 16; i += blockDim.x) reinterpret_cast<uint4 *>(b_ops)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int g = 0; g < NG; g++) {
            mbar_init(&bars[g * 5 + 0], 1);             // r committed
            mbar_init(&bars[g * 5 + 1], 1);             // z committed
            mbar_init(&bars[g * 5 + 2], 1);             // c committed
            mbar_init(&bars[g * 5 + 3], NQ);            // r*h operand written
            mbar_init(&bars[g * 5 + 4], NQ);            // h operand written
        }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, TCOLS);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    // ---- weights -> TMEM (once per layer): tiles r_hi r_lo z_hi z_lo c_hi c_lo -------------
    if (warp < 4) {
        const int m = tid;
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int g = 0; g < 3; g++) {
            // tile order r, z, c; the reference stores z rows first, then r (src/layers.c:511-526)
            const int mm = (m < H) ? m : 0;
            const float *row = (g == 0) ? (sW + (size_t)(H + mm) * H) : ((g == 1) ? (sW + (size_t)mm * H) : (sW2 + (size_t)mm * H));
#pragma unroll 1
            for (int kc = 0; kc < NKS; kc++) {
                uint32_t whi[8], wlo[8];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    float4 v = *reinterpret_cast<const float4 *>(row + kc * 16 + q * 4);
                    if (m >= H) v = make_float4(0.f, 0.f, 0.f, 0.f);
                    __half h0, l0, h1, l1, h2, l2, h3, l3;
                    split_fp16(v.x, h0, l0); split_fp16(v.y, h1, l1); split_fp16(v.z, h2, l2); split_fp16(v.w, h3, l3);
                    whi[2 * q] = pack_half2(h0, h1); whi[2 * q + 1] = pack_half2(h2, h3);
                    wlo[2 * q] = pack_half2(l0, l1); wlo[2 * q + 1] = pack_half2(l2, l3);
                }
                tmem_st8(lane_base + (2 * g) * KH + kc * 8, whi);
                tmem_st8(lane_base + (2 * g + 1) * KH + kc * 8, wlo);
            }
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // H <= 96: four warps per group -- gate warps 4g .. 4g+2 (lane quarters 0-2), issuer 4g+3, i.e. every issuer
    // sits on scheduler 3, which has no gate math; a CTA is 8 warps (two groups, 8 reads) or 16 warps (four
    // groups, 16 reads: the gate math of four groups interleaves on each scheduler and one SM carries twice the
    // reads).  H = 112: all eight warps 0-7 are gate warps, the issuers are warps 11 / 15 (512 threads).
    // H = 112 with three groups (12 reads, 336 + 144 TMEM columns): gate warps 0-11, issuers 13 / 14 / 15.
    const bool is_issuer = (NQ < 4) ? ((warp & 3) == 3 && warp < 4 * NG)
                                    : ((NG == 2) ? (warp == 11 || warp == 15) : (warp >= 13 && warp < 13 + NG));
    const bool is_gate = (warp < 4 * NG) && ((warp & 3) < NQ);
    const int grp = (NQ < 4) ? (warp >> 2) : (is_issuer ? ((NG == 2) ? (warp == 15) : (warp - 13)) : (warp >> 2));
    uint8_t *b_h = b_ops + grp * 2 * TILE_B, *b_rh = b_h + TILE_B;
    uint64_t *bar_r = &bars[grp * 5 + 0], *bar_z = &bars[grp * 5 + 1], *bar_c = &bars[grp * 5 + 2],
             *bar_rh = &bars[grp * 5 + 3], *bar_h = &bars[grp * 5 + 4];
    const uint32_t acc0 = tmem + ACC0 + grp * 3 * NM;
    int Tmax = 0;
    for (int i = 0; i < RPG; i++) {
        const int r = r0 + grp * RPG + i;
        if (r < d.nread) Tmax = max(Tmax, d.nblock[r]);
    }

    if (is_issuer) {
        // ---- UMMA issuer of one group -----------------------------------------------------------
        if (grp > 0) {                                   // stagger the groups over a step
            const long long t0 = clock64();
            const long long lag = (NG == 2) ? 700 : ((NG == 3) ? 650 : 450) * grp;
            while (clock64() - t0 < lag) { }
        }
        const uint32_t idesc = umma_idesc_f16(128, NM);
        const uint64_t dBh = umma_desc(smem_u32(b_h), LBO_B, SBO_B), dBrh = umma_desc(smem_u32(b_rh), LBO_B, SBO_B);
        constexpr uint64_t KB = (2 * LBO_B) >> 4;
        for (int s = 0; s < Tmax; s++) {
            if (s > 0) mbar_wait(bar_h, (s - 1) & 1);
            tc_fence_after();
            SB2_TRACE(0);
            if (elect_one()) {
#pragma unroll
                for (int g = 0; g < 2; g++) {
                    const uint32_t dcol = acc0 + g * NM;
                    const uint32_t w_hi = tmem + (2 * g) * KH, w_lo = w_hi + KH;
#pragma unroll
                    for (int ks = 0; ks < NKS; ks++) umma_f16_ts(dcol, w_lo + ks * 8, dBh + ks * KB, idesc, ks > 0);
#pragma unroll
                    for (int ks = 0; ks < NKS; ks++) umma_f16_ts(dcol, w_hi + ks * 8, dBh + ks * KB, idesc, 1);
                    umma_commit(g == 0 ? bar_r : bar_z);
                }
            }
            __syncwarp();
            SB2_TRACE(1);
            mbar_wait(bar_rh, s & 1);
            tc_fence_after();
            SB2_TRACE(2);
            if (elect_one()) {
                const uint32_t dcol = acc0 + 2 * NM;
                const uint32_t w_hi = tmem + 4 * KH, w_lo = w_hi + KH;
#pragma unroll
                for (int ks = 0; ks < NKS; ks++) umma_f16_ts(dcol, w_lo + ks * 8, dBrh + ks * KB, idesc, ks > 0);
#pragma unroll
                for (int ks = 0; ks < NKS; ks++) umma_f16_ts(dcol, w_hi + ks * 8, dBrh + ks * KB, idesc, 1);
                umma_commit(bar_c);
            }
            __syncwarp();
            SB2_TRACE(3);
        }
    } else if (is_gate) {
        // ---- gate warps ---------------------------------------------------------------------------
        const int q = warp & 3;
        const int j = q * 32 + lane;                    // hidden unit = accumulator row = TMEM lane
        const bool valid = j < H;
        const int jj = valid ? j : 0;
        const uint32_t acc_base = acc0 + ((uint32_t)(q * 32) << 16);
        int T[RPG];
        const float *xp[RPG];                           // this thread's element of the current input column
        const float *rp[RPG];
        float *op[RPG];
#pragma unroll
        for (int i = 0; i < RPG; i++) {
            const int r = r0 + grp * RPG + i;
            T[i] = (r < d.nread) ? d.nblock[r] : 0;
            const int col = (r < d.nread) ? d.col_off[r] : 0;
            const int t0 = backward ? max(T[i] - 1, 0) : 0;
            xp[i] = Xin + (size_t)(col + t0) * (3 * H) + jj;
            rp[i] = (resid != nullptr) ? resid + (size_t)(col + t0) * H + jj : nullptr;
            op[i] = out + (size_t)(col + t0) * H + jj;
        }
        const int xstep = backward ? -3 * H : 3 * H, ostep = backward ? -H : H;
        // operand element of (read i, unit j): row i (hi) / row 4 + i (lo) of k-group j / 8
        const uint32_t op_off = (uint32_t)(j >> 3) * LBO_B + (uint32_t)(j & 7) * 2;
        float h[RPG], xz[RPG], xr[RPG], xc[RPG], rs[RPG];
#pragma unroll
        for (int i = 0; i < RPG; i++) { h[i] = 0.0f; rs[i] = 0.0f; }
        auto load_x = [&](int s) {
#pragma unroll
            for (int i = 0; i < RPG; i++) {
                if (s < T[i]) {
                    xz[i] = xp[i][0]; xr[i] = xp[i][H]; xc[i] = xp[i][2 * H];
                    if (resid != nullptr) rs[i] = rp[i][0];
                } else {
                    xz[i] = 0.0f; xr[i] = 0.0f; xc[i] = 0.0f;
                }
            }
        };
        load_x(0);
        for (int s = 0; s < Tmax; s++) {
            float cz[RPG], cr[RPG], cc[RPG], crs[RPG];
#pragma unroll
            for (int i = 0; i < RPG; i++) { cz[i] = xz[i]; cr[i] = xr[i]; cc[i] = xc[i]; crs[i] = rs[i]; }
            float *ocur[RPG];
#pragma unroll
            for (int i = 0; i < RPG; i++) {
                ocur[i] = op[i];
                xp[i] += xstep; op[i] += ostep;
                if (resid != nullptr) rp[i] += ostep;
            }
            if (s + 1 < Tmax) load_x(s + 1);            // prefetch the next step's inputs

            // reset gate -> (r * h) operand
            mbar_wait(bar_r, s & 1);
            tc_fence_after();
            SB2_TRACE(4);
            {
                float a[8], pre[RPG], gr[RPG];
                tmem_ld8(acc_base, a);
                tmem_ld_wait();
                SB2_TRACE(5);
#pragma unroll
                for (int i = 0; i < RPG; i++) pre[i] = fmaf(a[i] + a[4 + i], RESULT_SCALE, cr[i]);
                sigmoid4<MATH>(pre, gr);
                float xs[RPG], fh[RPG];
                __half hi[RPG], lo[RPG];
#pragma unroll
                for (int i = 0; i < RPG; i++) xs[i] = gr[i] * h[i] * OPERAND_SCALE;
#pragma unroll
                for (int i = 0; i < RPG; i++) hi[i] = __float2half_rn(xs[i]);
#pragma unroll
                for (int i = 0; i < RPG; i++) fh[i] = __half2float(hi[i]);
#pragma unroll
                for (int i = 0; i < RPG; i++) lo[i] = __float2half_rn(xs[i] - fh[i]);
                if (valid) {
#pragma unroll
                    for (int i = 0; i < RPG; i++) {
                        *reinterpret_cast<__half *>(b_rh + op_off + i * 16) = hi[i];
                        *reinterpret_cast<__half *>(b_rh + op_off + (4 + i) * 16) = lo[i];
                    }
                }
            }
            SB2_TRACE(6);
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_rh);
            SB2_TRACE(7);

            // update gate (its UMMAs ran while the reset gate was being evaluated)
            float gz[RPG];
            mbar_wait(bar_z, s & 1);
            tc_fence_after();
            SB2_TRACE(8);
            {
                float a[8], pre[RPG];
                tmem_ld8(acc_base + NM, a);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < RPG; i++) pre[i] = fmaf(a[i] + a[4 + i], RESULT_SCALE, cz[i]);
                sigmoid4<MATH>(pre, gz);
            }
            SB2_TRACE(9);

            // candidate, state update, next step's operand
            mbar_wait(bar_c, s & 1);
            tc_fence_after();
            SB2_TRACE(10);
            float hn[RPG];
            {
                float a[8], pre[RPG], cand[RPG];
                tmem_ld8(acc_base + 2 * NM, a);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < RPG; i++) pre[i] = fmaf(a[i] + a[4 + i], RESULT_SCALE, cc[i]);
                tanh4<MATH>(pre, cand);
#pragma unroll
                for (int i = 0; i < RPG; i++) hn[i] = gz[i] * h[i] + (1.0f - gz[i]) * cand[i];
                float xs[RPG], fh[RPG];
                __half hi[RPG], lo[RPG];
#pragma unroll
                for (int i = 0; i < RPG; i++) { h[i] = hn[i]; xs[i] = hn[i] * OPERAND_SCALE; }
#pragma unroll
                for (int i = 0; i < RPG; i++) hi[i] = __float2half_rn(xs[i]);
#pragma unroll
                for (int i = 0; i < RPG; i++) fh[i] = __half2float(hi[i]);
#pragma unroll
                for (int i = 0; i < RPG; i++) lo[i] = __float2half_rn(xs[i] - fh[i]);
                if (valid) {
#pragma unroll
                    for (int i = 0; i < RPG; i++) {
                        *reinterpret_cast<__half *>(b_h + op_off + i * 16) = hi[i];
                        *reinterpret_cast<__half *>(b_h + op_off + (4 + i) * 16) = lo[i];
                    }
                }
            }
            SB2_TRACE(11);
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_h);
            SB2_TRACE(12);
            // results to HBM, off the critical path
            if (valid) {
#pragma unroll
                for (int i = 0; i < RPG; i++)
                    if (s < T[i]) *ocur[i] = (resid != nullptr) ? hn[i] + crs[i] : hn[i];
            }
        }
    }
#undef SB2_TRACE
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

template <int H, int MATH, int NG>
static int launch_scan_v4(const float *Xin, const float *sW, const float *sW2, const float *resid, float *out,
                          const BatchDims &d, int backward, long long *trace, cudaStream_t s) {
    // All 512 TMEM columns are allocated, so a second scan CTA on the same SM would stall in tcgen05.alloc: the
    // dynamic shared-memory request keeps it off (2 x (104 + 14) KB > 228 KB) while leaving ~110 KB for the
    // decode / conv CTAs of other batches that share the SM.
    constexpr int EXCLUSIVE_SMEM = 104 * 1024;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(gru_scan_v4_kernel<H, MATH, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 EXCLUSIVE_SMEM) != cudaSuccess)
            return -1;
        configured = true;
    }
    const int grid = (d.nread + 4 * NG - 1) / (4 * NG);
    gru_scan_v4_kernel<H, MATH, NG><<<grid, (H > 96) ? 512 : 128 * NG, EXCLUSIVE_SMEM, s>>>(Xin, sW, sW2, resid, out, d, backward, trace);
    return 0;
}

int launch_gru_scan_v4(const float *Xin, const float *sW, const float *sW2, const float *resid, float *out,
                       const BatchDims &d, int H, int backward, int math, long long *trace, cudaStream_t s) {
    // reads per CTA: 16 (four groups) once a batch has enough reads to fill the GPU twice over at 8 per CTA
    // (SCRAPPIE_B200_SCAN_GROUPS=2|4 overrides)
    static int groups = -1;
    if (groups < 0) { const char *e = getenv("SCRAPPIE_B200_SCAN_GROUPS"); groups = e ? atoi(e) : 0; }
    const bool four = (H == 96) && (groups == 4 || (groups == 0 && d.nread >= 128));
    const bool three = (H == 112) && (groups == 3 || (groups == 0 && d.nread >= 96));
#define SB2_CASE3(MM) if (three && math == MM) return launch_scan_v4<112, MM, 3>(Xin, sW, sW2, resid, out, d, backward, trace, s)
    SB2_CASE3(5); SB2_CASE3(2); SB2_CASE3(0);
#undef SB2_CASE3
#define SB2_CASE(HH, MM) if (H == HH && math == MM) return launch_scan_v4<HH, MM, 2>(Xin, sW, sW2, resid, out, d, backward, trace, s)
#define SB2_CASE4(MM) if (four && math == MM) return launch_scan_v4<96, MM, 4>(Xin, sW, sW2, resid, out, d, backward, trace, s)
    SB2_CASE4(5); SB2_CASE4(2); SB2_CASE4(0);
    SB2_CASE(96, 0); SB2_CASE(96, 1); SB2_CASE(96, 2);
    SB2_CASE(112, 0); SB2_CASE(112, 1); SB2_CASE(112, 2);
    SB2_CASE(96, 3); SB2_CASE(112, 3); SB2_CASE(96, 4); SB2_CASE(112, 4); SB2_CASE(96, 5); SB2_CASE(112, 5);
#undef SB2_CASE
#undef SB2_CASE4
    return -1;
}

template <int H, int NR>
static size_t scan_smem_bytes() {
    constexpr int NM = (NR < 16) ? 16 : NR;
    return (size_t)ScanLayout<H>::WEIGHTS + 4 * (size_t)(H / 8) * (16 * NM + 16) + (size_t)((128 - H) / 8) * ScanLayout<H>::SBO_A + 64;
}

template <int H, int NR, bool FAST>
static int launch_scan(const float *Xin, const uint8_t *wimg, const float *resid, float *out, const BatchDims &d,
                       int backward, cudaStream_t s) {
    const size_t smem = scan_smem_bytes<H, NR>();
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(gru_scan_tc_kernel<H, NR, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return -1;
        configured = true;
    }
    const int grid = (d.nread + NR - 1) / NR;
    gru_scan_tc_kernel<H, NR, FAST><<<grid, 160, smem, s>>>(Xin, wimg, resid, out, d, backward);
    return 0;
}

int launch_gru_scan_tc(const float *Xin, const uint8_t *wimg, const float *resid, float *out, const BatchDims &d,
                       int H, int backward, int fast_math, cudaStream_t s) {
    if (H == 96) return fast_math ? launch_scan<96, 8, true>(Xin, wimg, resid, out, d, backward, s)
                                  : launch_scan<96, 8, false>(Xin, wimg, resid, out, d, backward, s);
    if (H == 112) return fast_math ? launch_scan<112, 8, true>(Xin, wimg, resid, out, d, backward, s)
                                   : launch_scan<112, 8, false>(Xin, wimg, resid, out, d, backward, s);
    return -1;
}

}  // namespace sb2
