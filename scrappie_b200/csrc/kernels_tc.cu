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

__device__ __forceinline__ uint32_t pack_half2(__half lo16, __half hi16) {
    return (uint32_t)__half_as_ushort(lo16) | ((uint32_t)__half_as_ushort(hi16) << 16);
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
                for (int i = 0; i < RPG; i++) hn[i] = __fmaf_rn(gz[i], h[i], __fmul_rn(1.0f - gz[i], cand[i]));   // pinned: v4 and v5 must round alike
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

// ---------------------------------------------------------------------------------
// GRU scan v5: eight reads per group (the UMMA's N = 16 fully used), inputs through a TMA ring
// ---------------------------------------------------------------------------------
// v4 spends its step on the CUDA-core side: per thread 4 reads x 1 hidden unit, of which a large part is
// overhead paid per step rather than per read -- three mbarrier waits, three tcgen05.ld + wait, two operand
// hand-overs (fence, arrive), 12 scalar global loads with 64-bit pointer bumps -- and its B operand uses only 8
// of the 16 rows the instruction computes anyway.  v5 keeps v4's data flow (weights in TMEM, A from tensor
// memory, hi | lo of the state side by side in N, reset gate committed first, one issuer warp per group on
// scheduler 3) and changes the shape of the work:
//   * a group is EIGHT reads: rows 0..7 of the B operand hold fp16 hi, rows 8..15 fp16 lo; the same 36
//     tcgen05.mma (M128 N16 K16) per group-step now serve twice the reads (NG = 4 groups: 32 reads per CTA at
//     H = 96, 288 + 4 x 48 = 480 TMEM columns; NG = 3: 24 reads at H = 112, 336 + 144 columns);
//   * the per-step input columns Xin[t] (3H floats per read, contiguous) -- and the residual input for rnnrf --
//     arrive through a shared-memory ring filled by cp.async.bulk (TMA, SASS UBLKCP), issued by the group's
//     issuer warp two steps ahead: lane i < 8 copies read i's column, completion is counted in bytes on the
//     slot's mbarrier.  No register holds a load in flight, no 64-bit pointer arithmetic per step; the slot a
//     copy overwrites was last read before the gate warps' bar_h arrive that the issuer has just waited for;
//   * the gate math is written for 8 independent chains per thread (sigmoid_n / tanh_n).
constexpr int V5_RING = 3;              // slots: the copy for step s + 2 is issued while step s runs

template <int MATH, int N>
__device__ __forceinline__ void sigmoid_n(const float (&x)[N], float (&y)[N]) {
    if (MATH == 0) {
#pragma unroll
        for (int i = 0; i < N; i++) y[i] = logistic_cephes(x[i]);
    } else if (MATH == 5) {
        // ex2.approx on the plainly rounded argument, Newton-refined reciprocal
        float e[N], dd[N], q[N];
#pragma unroll
        for (int i = 0; i < N; i++) e[i] = ex2_approx(fminf(x[i] * -1.4426950408889634f, 126.0f));
#pragma unroll
        for (int i = 0; i < N; i++) dd[i] = 1.0f + e[i];
#pragma unroll
        for (int i = 0; i < N; i++) q[i] = rcp_approx(dd[i]);
#pragma unroll
        for (int i = 0; i < N; i++) y[i] = fmaf(q[i], fmaf(-dd[i], q[i], 1.0f), q[i]);
    } else {
        // degree-6 polynomial 2^f on [-1/2, 1/2], exponent by the magic-number trick, Newton-refined reciprocal
        float t[N], r[N], f[N], p[N], dd[N], q[N];
#pragma unroll
        for (int i = 0; i < N; i++) t[i] = fmaxf(fminf(x[i] * -1.4426950408889634f, 126.0f), -126.0f);
#pragma unroll
        for (int i = 0; i < N; i++) r[i] = t[i] + 12582912.0f;
#pragma unroll
        for (int i = 0; i < N; i++) f[i] = t[i] - (r[i] - 12582912.0f);
#pragma unroll
        for (int i = 0; i < N; i++) p[i] = fmaf(0.00015337577497120947f, f[i], 0.0013399859890341759f);
#pragma unroll
        for (int i = 0; i < N; i++) p[i] = fmaf(p[i], f[i], 0.009618519805371761f);
#pragma unroll
        for (int i = 0; i < N; i++) p[i] = fmaf(p[i], f[i], 0.05550329014658928f);
#pragma unroll
        for (int i = 0; i < N; i++) p[i] = fmaf(p[i], f[i], 0.24022646248340607f);
#pragma unroll
        for (int i = 0; i < N; i++) p[i] = fmaf(p[i], f[i], 0.6931471824645996f);
#pragma unroll
        for (int i = 0; i < N; i++) p[i] = fmaf(p[i], f[i], 1.0f);
#pragma unroll
        for (int i = 0; i < N; i++)
            dd[i] = 1.0f + __int_as_float(__float_as_int(p[i]) + ((__float_as_int(r[i]) - 0x4B400000) << 23));
#pragma unroll
        for (int i = 0; i < N; i++) q[i] = rcp_approx(dd[i]);
#pragma unroll
        for (int i = 0; i < N; i++) y[i] = fmaf(q[i], fmaf(-dd[i], q[i], 1.0f), q[i]);
    }
}
template <int MATH, int N>
__device__ __forceinline__ void tanh_n(const float (&x)[N], float (&y)[N]) {
    if (MATH == 0) {
#pragma unroll
        for (int i = 0; i < N; i++) y[i] = tanh_cephes(x[i]);
    } else {
        float x2[N], sg[N];
#pragma unroll
        for (int i = 0; i < N; i++) x2[i] = x[i] + x[i];
        sigmoid_n<MATH, N>(x2, sg);
#pragma unroll
        for (int i = 0; i < N; i++) y[i] = (sg[i] + sg[i]) - 1.0f;
    }
}

template <int H, int NG, bool RESID>
struct ScanV5Cfg {
    static constexpr int RPG = 8, NM = 16;
    static constexpr uint32_t LBO_B = 16u * NM + 16u, SBO_B = 128u;
    static constexpr uint32_t TILE_B = (H / 8) * LBO_B;
    static constexpr uint32_t XCOL_B = 3 * H * 4, RCOL_B = RESID ? H * 4 : 0;     // bytes per read and step
    static constexpr uint32_t SLOT_B = RPG * (XCOL_B + RCOL_B);                    // one group, one step
    static constexpr uint32_t OFF_OPS = 0;                                        // per group: h [hi|lo], r*h [hi|lo]
    static constexpr uint32_t OFF_RING = (NG * 2 * TILE_B + 127) / 128 * 128;
    static constexpr uint32_t OFF_BAR = OFF_RING + NG * V5_RING * SLOT_B;
    static constexpr uint32_t NBAR = NG * (5 + V5_RING);
    static constexpr uint32_t OFF_META = OFF_BAR + NBAR * 8 + 16;                  // per group: cbase[8], T[8]
    static constexpr uint32_t SMEM = OFF_META + NG * 16 * 4;
    // all 512 TMEM columns are allocated: a second scan CTA on the SM would stall in tcgen05.alloc, so the request
    // is at least half of the SM's shared memory
    static constexpr uint32_t SMEM_REQ = SMEM > 116 * 1024 ? SMEM : 116 * 1024;
    static_assert(SLOT_B % 16 == 0 && XCOL_B % 16 == 0 && RCOL_B % 16 == 0, "bulk copy alignment");
    static_assert(SMEM <= 200 * 1024, "shared memory budget");
};

template <int H, int MATH, int NG, bool RESID>
__global__ void __launch_bounds__(512, 1)
gru_scan_v5_kernel(const float *__restrict__ Xin, const float *__restrict__ sW, const float *__restrict__ sW2,
                   const float *__restrict__ resid, float *__restrict__ out, BatchDims d, int backward) {
    using C = ScanV5Cfg<H, NG, RESID>;
    constexpr int RPG = C::RPG, NM = C::NM;
    static_assert((NG == 4 && H <= 96) || (NG == 3 && H > 96), "groups per CTA: TMEM columns and warp slots");
    constexpr uint32_t LBO_B = C::LBO_B, SBO_B = C::SBO_B, TILE_B = C::TILE_B;
    constexpr int NKS = H / 16;
    constexpr int NQ = (H + 31) / 32;                   // lane quarters that own hidden units
    constexpr uint32_t KH = H / 2;                      // TMEM columns per weight tile
    constexpr uint32_t ACC0 = 6 * KH;                   // accumulators: per group r, z, c (16 columns each)
    constexpr uint32_t TCOLS = 512;
    static_assert(ACC0 + NG * 3 * NM <= TCOLS, "TMEM budget");

    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *b_ops = smem + C::OFF_OPS;
    uint8_t *ring = smem + C::OFF_RING;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + C::OFF_BAR + C::NBAR * 8);
    int *meta = reinterpret_cast<int *>(smem + C::OFF_META);

    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int r0 = blockIdx.x * (NG * RPG);

    for (uint32_t i = tid; i < NG * 2 * TILE_B / 16; i += blockDim.x) reinterpret_cast<uint4 *>(b_ops)[i] = make_uint4(0, 0, 0, 0);
    if (tid < NG * RPG) {
        // first column this read touches (t = 0 forward, T - 1 backward) and its length
        const int r = r0 + tid;
        const int T = (r < d.nread) ? d.nblock[r] : 0;
        const int col = (r < d.nread) ? d.col_off[r] : 0;
        meta[(tid / RPG) * 16 + (tid % RPG)] = backward ? col + max(T - 1, 0) : col;
        meta[(tid / RPG) * 16 + 8 + (tid % RPG)] = T;
    }
    if (tid == 0) {
        for (int g = 0; g < NG; g++) {
            uint64_t *gb = bars + g * (5 + V5_RING);
            mbar_init(&gb[0], 1);                       // r committed
            mbar_init(&gb[1], 1);                       // z committed
            mbar_init(&gb[2], 1);                       // c committed
            mbar_init(&gb[3], NQ);                      // r*h operand written
            mbar_init(&gb[4], NQ);                      // h operand written
            for (int k = 0; k < V5_RING; k++) mbar_init(&gb[5 + k], 1);     // input slot k filled (transaction bytes)
        }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, TCOLS);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

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

    // H <= 96: four warps per group -- gate warps 4g .. 4g+2 (TMEM lane quarters 0-2), issuer 4g+3, so every issuer
    // sits on scheduler 3, which has no gate math.  H = 112: gate warps 0-11 (four quarters per group), issuers 13-15.
    const bool is_issuer = (NQ < 4) ? ((warp & 3) == 3 && warp < 4 * NG) : (warp >= 13 && warp < 13 + NG);
    const bool is_gate = (warp < 4 * NG) && ((warp & 3) < NQ);
    const int grp = (NQ < 4) ? (warp >> 2) : (is_issuer ? (warp - 13) : (warp >> 2));
    if (!is_issuer && !is_gate) {
        // idle warps only take part in the final barrier
    }
    const int gsel = (is_issuer || is_gate) ? grp : 0;
    uint8_t *b_h = b_ops + gsel * 2 * TILE_B, *b_rh = b_h + TILE_B;
    uint64_t *gb = bars + gsel * (5 + V5_RING);
    uint64_t *bar_r = &gb[0], *bar_z = &gb[1], *bar_c = &gb[2], *bar_rh = &gb[3], *bar_h = &gb[4], *bar_x = &gb[5];
    uint8_t *gring = ring + (size_t)gsel * V5_RING * C::SLOT_B;
    const int *gmeta = meta + gsel * 16;
    const uint32_t acc0 = tmem + ACC0 + gsel * 3 * NM;
    int Tmax = 0;
#pragma unroll
    for (int i = 0; i < RPG; i++) Tmax = max(Tmax, gmeta[8 + i]);
    const int dir = backward ? -1 : 1;

    if (is_issuer) {
        // ---- UMMA issuer + input-ring producer of one group ------------------------------------------
        // lane i < 8 owns read i of the group: its column pointer advances by one column per step
        const int myT = (lane < RPG) ? gmeta[8 + lane] : 0;
        const float *xsrc = Xin + (size_t)((lane < RPG) ? gmeta[lane] : 0) * (3 * H);
        const float *rsrc = RESID ? (resid + (size_t)((lane < RPG) ? gmeta[lane] : 0) * H) : nullptr;
        auto fill = [&](int st) {                       // request the inputs of step st (all lanes call it)
            if (st < Tmax) {
                const int slot = st % V5_RING;
                const bool mine = st < myT;             // lanes >= 8 have myT = 0
                const unsigned vm = __ballot_sync(0xffffffffu, mine);
                if (lane == 0) mbar_arrive_expect_tx(&bar_x[slot], (uint32_t)__popc(vm) * (C::XCOL_B + C::RCOL_B));
                __syncwarp();
                if (mine) {
                    uint8_t *dst = gring + slot * C::SLOT_B + lane * (C::XCOL_B + C::RCOL_B);
                    bulk_g2s(dst, xsrc + (ptrdiff_t)st * dir * (3 * H), C::XCOL_B, &bar_x[slot]);
                    if (RESID) bulk_g2s(dst + C::XCOL_B, rsrc + (ptrdiff_t)st * dir * H, C::RCOL_B, &bar_x[slot]);
                }
            }
        };
#pragma unroll 1
        for (int st = 0; st < V5_RING - 1; st++) fill(st);
        if (grp > 0) {                                   // stagger the groups over a step
            const long long t0 = clock64();
            const long long lag = ((NG == 3) ? 900 : 650) * grp;
            while (clock64() - t0 < lag) { }
        }
        const uint32_t idesc = umma_idesc_f16(128, NM);
        const uint64_t dBh = umma_desc(smem_u32(b_h), LBO_B, SBO_B), dBrh = umma_desc(smem_u32(b_rh), LBO_B, SBO_B);
        constexpr uint64_t KB = (2 * LBO_B) >> 4;
        for (int s = 0; s < Tmax; s++) {
            if (s > 0) mbar_wait(bar_h, (s - 1) & 1);
            tc_fence_after();
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
            // every gate warp has finished step s - 1 (bar_h), so the slot step s - 1 used is free again
            fill(s + V5_RING - 1);
            mbar_wait(bar_rh, s & 1);
            tc_fence_after();
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
        }
    } else if (is_gate) {
        // ---- gate warps ---------------------------------------------------------------------------
        const int q = warp & 3;
        const int j = q * 32 + lane;                    // hidden unit = accumulator row = TMEM lane
        const bool valid = j < H;
        const int jj = valid ? j : 0;
        const uint32_t acc_base = acc0 + ((uint32_t)(q * 32) << 16);
        // operand element of (read i, unit j): row i (hi) / row 8 + i (lo, the second 8-row group) of k-group j / 8
        const uint32_t op_off = (uint32_t)(j >> 3) * LBO_B + (uint32_t)(j & 7) * 2;
        int T[RPG], ocol[RPG];
#pragma unroll
        for (int i = 0; i < RPG; i++) { ocol[i] = gmeta[i]; T[i] = gmeta[8 + i]; }
        float h[RPG];
#pragma unroll
        for (int i = 0; i < RPG; i++) h[i] = 0.0f;
        for (int s = 0; s < Tmax; s++) {
            const int slot = s % V5_RING;
            const float *xs_ = reinterpret_cast<const float *>(gring + slot * C::SLOT_B) + jj;
            constexpr int XSTR = (C::XCOL_B + C::RCOL_B) / 4;     // floats between consecutive reads of a slot
            mbar_wait(&bar_x[slot], (s / V5_RING) & 1);            // this step's input columns have landed

            // reset gate -> (r * h) operand
            mbar_wait(bar_r, s & 1);
            tc_fence_after();
            {
                float a[16], pre[RPG], gr[RPG];
                tmem_ld16(acc_base, a);
#pragma unroll
                for (int i = 0; i < RPG; i++) pre[i] = (s < T[i]) ? xs_[i * XSTR + H] : 0.0f;
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < RPG; i++) pre[i] = fmaf(a[i] + a[8 + i], RESULT_SCALE, pre[i]);
                sigmoid_n<MATH, RPG>(pre, gr);
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
                        *reinterpret_cast<__half *>(b_rh + op_off + SBO_B + i * 16) = lo[i];
                    }
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_rh);

            // update gate (its UMMAs ran while the reset gate was being evaluated)
            float gz[RPG];
            mbar_wait(bar_z, s & 1);
            tc_fence_after();
            {
                float a[16], pre[RPG];
                tmem_ld16(acc_base + NM, a);
#pragma unroll
                for (int i = 0; i < RPG; i++) pre[i] = (s < T[i]) ? xs_[i * XSTR] : 0.0f;
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < RPG; i++) pre[i] = fmaf(a[i] + a[8 + i], RESULT_SCALE, pre[i]);
                sigmoid_n<MATH, RPG>(pre, gz);
            }

            // candidate, state update, next step's operand
            float xc[RPG], rs[RPG];
#pragma unroll
            for (int i = 0; i < RPG; i++) {
                xc[i] = (s < T[i]) ? xs_[i * XSTR + 2 * H] : 0.0f;
                rs[i] = (RESID && s < T[i]) ? xs_[i * XSTR + 3 * H] : 0.0f;
            }
            mbar_wait(bar_c, s & 1);
            tc_fence_after();
            float hn[RPG];
            {
                float a[16], pre[RPG], cand[RPG];
                tmem_ld16(acc_base + 2 * NM, a);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < RPG; i++) pre[i] = fmaf(a[i] + a[8 + i], RESULT_SCALE, xc[i]);
                tanh_n<MATH, RPG>(pre, cand);
#pragma unroll
                for (int i = 0; i < RPG; i++) hn[i] = __fmaf_rn(gz[i], h[i], __fmul_rn(1.0f - gz[i], cand[i]));   // pinned: v4 and v5 must round alike
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
                        *reinterpret_cast<__half *>(b_h + op_off + SBO_B + i * 16) = lo[i];
                    }
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_h);
            // results to HBM, off the critical path
            if (valid) {
#pragma unroll
                for (int i = 0; i < RPG; i++)
                    if (s < T[i]) out[(size_t)(ocol[i] + s * dir) * H + j] = RESID ? hn[i] + rs[i] : hn[i];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

// ---------------------------------------------------------------------------------
// GRU scan v6: v5's data flow, packed fp32 gate math, results through a TMA store
// ---------------------------------------------------------------------------------
// v5's gate warps issue ~580 instructions per step for 8 reads (profiles/r2a): 72 % of the three gate schedulers'
// issue slots, next to a MUFU pipe that is half busy -- the kernel is bound by CUDA-core issue.  v6 cuts the count:
//   * fp32 arithmetic on PAIRS of reads with the packed instructions sm_100 has (fma / add / mul .f32x2 -> FFMA2,
//     FADD2, FMUL2): the eight reads of a thread are four register pairs;
//   * the state is kept pre-scaled by 2^8 (the operand scale), tanh's doubling is folded into its exp2 constant, the
//     blend is two packed FMAs;
//   * no per-read predication of the input loads (a finished read's rows of the operand only feed its own,
//     never stored, accumulator columns) and no address arithmetic for the results: the gate warps put a step's
//     output into a shared-memory staging row and the group's issuer warp stores it with cp.async.bulk (one 4 H-byte
//     row per read), after the UMMAs of the next step are on their way;
//   * RPG = 4 or 8 reads per group from the same code (small batches take 4: shorter steps), so the two
//     configurations round identically and a read's result does not depend on the batch it travels in.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 splat2(float a) { return pk2(a, a); }

// shared memory -> global bulk copy (TMA store), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// MATH 5: y = 1 / (1 + 2^t) for NP pairs, t = -x log2(e): ex2.approx on the clamped argument, rcp.approx refined by one
// Newton step (the refinement, not the exponential, is what the 1e-5 probability-space bound needs: DESIGN.md section 3).
// MATH 0: y = logistic(t) by the reference's cephes polynomial, element by element (t is the pre-activation itself).
template <int MATH, int NP>
__device__ __forceinline__ void logistic_pk(const f32x2 (&t)[NP], f32x2 (&y)[NP]) {
    const f32x2 neg1 = splat2(-1.0f), one = splat2(1.0f);
#pragma unroll
    for (int p = 0; p < NP; p++) {
        float t0, t1;
        upk2(t[p], t0, t1);
        if (MATH == 0) {
            y[p] = pk2(logistic_cephes(t0), logistic_cephes(t1));            // MATH 0: the caller passes x itself
        } else {
            const float e0 = ex2_approx(fminf(t0, 126.0f)), e1 = ex2_approx(fminf(t1, 126.0f));
            const f32x2 dn = fma2(pk2(e0, e1), neg1, neg1);                 // -(1 + e)
            float d0, d1;
            upk2(dn, d0, d1);
            const f32x2 q = pk2(rcp_approx(-d0), rcp_approx(-d1));
            y[p] = fma2(q, fma2(dn, q, one), q);                            // q + q (1 - d q)
        }
    }
}

template <int H, int NG, int RPG, bool RESID>
struct ScanV6Cfg {
    static constexpr int NM = 16;
    static constexpr int NQ = (H + 31) / 32;
    static constexpr int NTHREADS = (NQ < 4) ? 128 * NG : 32 * (5 * NG + 1);     // H = 112: gate warps 0 .. 4 NG - 1, issuers 4 NG + 1 ..
    static constexpr uint32_t LBO_B = 16u * NM + 16u, SBO_B = 128u;
    static constexpr uint32_t TILE_B = (H / 8) * LBO_B;
    static constexpr uint32_t XCOL_B = 3 * H * 4, RCOL_B = RESID ? H * 4 : 0;     // input bytes per read and step
    static constexpr uint32_t SLOT_B = RPG * (XCOL_B + RCOL_B);                    // one group, one step
    static constexpr uint32_t OUT_B = RPG * H * 4;                                 // one group, one step of results
    static constexpr uint32_t OFF_RING = (NG * 2 * TILE_B + 127) / 128 * 128;
    static constexpr uint32_t OFF_OUT = OFF_RING + NG * V5_RING * SLOT_B;
    static constexpr uint32_t OFF_BAR = OFF_OUT + NG * 2 * OUT_B;
    static constexpr uint32_t NBAR = NG * (5 + V5_RING);
    static constexpr uint32_t OFF_META = OFF_BAR + NBAR * 8 + 16;                  // per group: first column [8], length [8]
    static constexpr uint32_t SMEM = OFF_META + NG * 16 * 4;
    // all 512 TMEM columns are allocated: a second scan CTA on the SM would stall in tcgen05.alloc, so the request
    // is at least half of the SM's shared memory
    static constexpr uint32_t SMEM_REQ = SMEM > 116 * 1024 ? SMEM : 116 * 1024;
    static_assert(SLOT_B % 16 == 0 && XCOL_B % 16 == 0 && RCOL_B % 16 == 0 && OUT_B % 16 == 0, "bulk copy alignment");
    static_assert(SMEM <= 200 * 1024, "shared memory budget");
};

template <int H, int MATH, int NG, int RPG, bool RESID>
__global__ void __launch_bounds__(ScanV6Cfg<H, NG, RPG, RESID>::NTHREADS, 1)
gru_scan_v6_kernel(const float *__restrict__ Xin, const float *__restrict__ sW, const float *__restrict__ sW2,
                   const float *__restrict__ resid, float *__restrict__ out, BatchDims d, int backward,
                   long long *__restrict__ trace) {
    using C = ScanV6Cfg<H, NG, RPG, RESID>;
    // diagnostic (SCRAPPIE_B200_TRACE=1): clock64() of CTA 0 at the hand-over points of steps 100..103, per group:
    // trace[(grp * 4 + step - 100) * 16 + slot]; slots 0-4 issuer, 5-11 gate warp of lane quarter 0
#define V6_TRACE(slot) do { if (trace != nullptr && blockIdx.x == 0 && lane == 0 && s >= 100 && s < 104) trace[((grp * 4) + (s - 100)) * 16 + (slot)] = clock64(); } while (0)
    constexpr int NM = C::NM, NP = RPG / 2, NQ = C::NQ;
    static_assert(RPG == 4 || RPG == 8, "reads per group");
    constexpr uint32_t LBO_B = C::LBO_B, SBO_B = C::SBO_B, TILE_B = C::TILE_B;
    constexpr int NKS = H / 16;
    constexpr uint32_t KH = H / 2;                      // TMEM columns per weight tile
    constexpr uint32_t ACC0 = 6 * KH;                   // accumulators: per group r, z, c (16 columns each)
    constexpr uint32_t TCOLS = 512;
    static_assert(ACC0 + NG * 3 * NM <= TCOLS, "TMEM budget");

    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *b_ops = smem;
    uint8_t *ring = smem + C::OFF_RING;
    uint8_t *ostage = smem + C::OFF_OUT;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + C::OFF_BAR + C::NBAR * 8);
    int *meta = reinterpret_cast<int *>(smem + C::OFF_META);

    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int r0 = blockIdx.x * (NG * RPG);

    // operands and input ring start as zeros: rows of reads that do not exist must stay finite-free of surprises
    for (uint32_t i = tid; i < C::OFF_OUT / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid < NG * RPG) {
        // first column this read touches (t = 0 forward, T - 1 backward) and its length
        const int r = r0 + tid;
        const int T = (r < d.nread) ? d.nblock[r] : 0;
        const int col = (r < d.nread) ? d.col_off[r] : 0;
        meta[(tid / RPG) * 16 + (tid % RPG)] = backward ? col + max(T - 1, 0) : col;
        meta[(tid / RPG) * 16 + 8 + (tid % RPG)] = T;
    }
    if (tid == 0) {
        for (int g = 0; g < NG; g++) {
            uint64_t *gb = bars + g * (5 + V5_RING);
            mbar_init(&gb[0], 1);                       // r committed
            mbar_init(&gb[1], 1);                       // z committed
            mbar_init(&gb[2], 1);                       // c committed
            mbar_init(&gb[3], NQ);                      // r*h operand written
            mbar_init(&gb[4], NQ);                      // h operand + result row written
            for (int k = 0; k < V5_RING; k++) mbar_init(&gb[5 + k], 1);     // input slot k filled (transaction bytes)
        }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, TCOLS);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

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

    // H <= 96: four warps per group -- gate warps 4g .. 4g+2 (TMEM lane quarters 0-2), issuer 4g+3: every issuer sits
    // on scheduler 3, which has no gate math.  H = 112: gate warps 0 .. 4 NG - 1 (four quarters per group), issuers
    // 4 NG + 1 + g (schedulers 1-3).
    const bool is_issuer = (NQ < 4) ? ((warp & 3) == 3 && warp < 4 * NG) : (warp > 4 * NG && warp <= 5 * NG);
    const bool is_gate = (warp < 4 * NG) && ((warp & 3) < NQ);
    const int grp = (NQ < 4) ? (warp >> 2) : (is_issuer ? (warp - 4 * NG - 1) : (warp >> 2));
    const int gsel = (is_issuer || is_gate) ? grp : 0;
    uint8_t *b_h = b_ops + gsel * 2 * TILE_B, *b_rh = b_h + TILE_B;
    uint64_t *gb = bars + gsel * (5 + V5_RING);
    uint64_t *bar_r = &gb[0], *bar_z = &gb[1], *bar_c = &gb[2], *bar_rh = &gb[3], *bar_h = &gb[4], *bar_x = &gb[5];
    uint8_t *gring = ring + (size_t)gsel * V5_RING * C::SLOT_B;
    uint8_t *gout = ostage + (size_t)gsel * 2 * C::OUT_B;
    const int *gmeta = meta + gsel * 16;
    const uint32_t acc0 = tmem + ACC0 + gsel * 3 * NM;
    int Tmax = 0;
#pragma unroll
    for (int i = 0; i < RPG; i++) Tmax = max(Tmax, gmeta[8 + i]);
    const int dir = backward ? -1 : 1;

    if (is_issuer) {
        // ---- UMMA issuer, input-ring producer and result writer of one group --------------------------
        // lane i < RPG owns read i of the group
        const int myT = (lane < RPG) ? gmeta[8 + lane] : 0;
        const int mycol = (lane < RPG) ? gmeta[lane] : 0;
        const float *xsrc = Xin + (size_t)mycol * (3 * H);
        const float *rsrc = RESID ? (resid + (size_t)mycol * H) : nullptr;
        float *odst = out + (size_t)mycol * H;
        auto fill = [&](int st) {                       // request the inputs of step st (all lanes call it)
            if (st < Tmax) {
                const int slot = st % V5_RING;
                const bool mine = st < myT;             // lanes >= RPG have myT = 0
                const unsigned vm = __ballot_sync(0xffffffffu, mine);
                if (lane == 0) mbar_arrive_expect_tx(&bar_x[slot], (uint32_t)__popc(vm) * (C::XCOL_B + C::RCOL_B));
                __syncwarp();
                if (mine) {
                    uint8_t *dst = gring + slot * C::SLOT_B + lane * (C::XCOL_B + C::RCOL_B);
                    bulk_g2s(dst, xsrc + (ptrdiff_t)st * dir * (3 * H), C::XCOL_B, &bar_x[slot]);
                    if (RESID) bulk_g2s(dst + C::XCOL_B, rsrc + (ptrdiff_t)st * dir * H, C::RCOL_B, &bar_x[slot]);
                }
            }
        };
        auto store = [&](int st) {                      // write the results of step st (staged by the gate warps)
            if (st < myT) bulk_s2g(odst + (ptrdiff_t)st * dir * H, gout + (st & 1) * C::OUT_B + lane * (H * 4), H * 4);
            bulk_commit();
        };
#pragma unroll 1
        for (int st = 0; st < V5_RING - 1; st++) fill(st);
        if (grp > 0) {                                   // stagger the groups over a step
            const long long t0 = clock64();
            const long long lag = (long long)(RPG == 8 ? 450 : 350) * grp;
            while (clock64() - t0 < lag) { }
        }
        const uint32_t idesc = umma_idesc_f16(128, NM);
        const uint64_t dBh = umma_desc(smem_u32(b_h), LBO_B, SBO_B), dBrh = umma_desc(smem_u32(b_rh), LBO_B, SBO_B);
        constexpr uint64_t KB = (2 * LBO_B) >> 4;
        for (int s = 0; s < Tmax; s++) {
            if (s > 0) mbar_wait(bar_h, (s - 1) & 1);
            tc_fence_after();
            V6_TRACE(0);
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
            V6_TRACE(1);
            // Every gate warp has finished step s - 1 (bar_h): its result row is staged and the input slot it used is
            // free.  The staging row step s will overwrite was last read by the store issued one step ago.
            bulk_wait_read0();
            if (s > 0) store(s - 1);
            fill(s + V5_RING - 1);
            V6_TRACE(2);
            mbar_wait(bar_rh, s & 1);
            tc_fence_after();
            V6_TRACE(3);
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
            V6_TRACE(4);
        }
        if (Tmax > 0) {
            mbar_wait(bar_h, (Tmax - 1) & 1);
            store(Tmax - 1);
        }
        bulk_wait0();                                   // results are in global memory before the CTA retires
    } else if (is_gate) {
        // ---- gate warps ---------------------------------------------------------------------------
        const int q = warp & 3;
        const int j = q * 32 + lane;                    // hidden unit = accumulator row = TMEM lane
        const bool valid = j < H;
        const int jj = valid ? j : 0;
        const uint32_t acc_base = acc0 + ((uint32_t)(q * 32) << 16);
        // operand element of (read i, unit j): row i (hi) / row 8 + i (lo, the second 8-row group) of k-group j / 8
        const uint32_t op_off = (uint32_t)(j >> 3) * LBO_B + (uint32_t)(j & 7) * 2;
        constexpr int XSTR = (C::XCOL_B + C::RCOL_B) / 4;       // floats between consecutive reads of an input slot
        const f32x2 rscale = splat2(RESULT_SCALE);
        // exponent-argument constants; the cephes mirror (MATH 0) takes the pre-activation itself
        const f32x2 k_sig = splat2(MATH == 0 ? 1.0f : -1.4426950408889634f), k_tanh = splat2(MATH == 0 ? 1.0f : -2.8853900817779268f);
        f32x2 hs[NP];                                   // state scaled by 2^8 (the operand scale): exact, and the form both uses want
#pragma unroll
        for (int p = 0; p < NP; p++) hs[p] = splat2(0.0f);

        // accumulator columns i (W h_hi) and 8 + i (W h_lo) of gate `g`, plus this step's input: -> exponent argument
        auto preact = [&](uint32_t col, const float *xcol, f32x2 kexp, f32x2 (&t)[NP]) {
            float a[16];
            tmem_ld16(acc_base + col, a);
            f32x2 x[NP];
#pragma unroll
            for (int p = 0; p < NP; p++) x[p] = pk2(xcol[(2 * p) * XSTR], xcol[(2 * p + 1) * XSTR]);
            tmem_ld_wait();
#pragma unroll
            for (int p = 0; p < NP; p++) {
                const f32x2 sum = add2(pk2(a[2 * p], a[2 * p + 1]), pk2(a[8 + 2 * p], a[8 + 2 * p + 1]));
                t[p] = mul2(fma2(sum, rscale, x[p]), kexp);
            }
        };
        // fp16 hi / lo rows of the operand `dst` from values already scaled by 2^8
        auto write_operand = [&](uint8_t *dst, const f32x2 (&v)[NP]) {
#pragma unroll
            for (int p = 0; p < NP; p++) {
                float v0, v1;
                upk2(v[p], v0, v1);
                const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
                const __half l0 = __float2half_rn(v0 - __half2float(h0)), l1 = __float2half_rn(v1 - __half2float(h1));
                if (valid) {
                    *reinterpret_cast<__half *>(dst + op_off + (2 * p) * 16) = h0;
                    *reinterpret_cast<__half *>(dst + op_off + (2 * p + 1) * 16) = h1;
                    *reinterpret_cast<__half *>(dst + op_off + SBO_B + (2 * p) * 16) = l0;
                    *reinterpret_cast<__half *>(dst + op_off + SBO_B + (2 * p + 1) * 16) = l1;
                }
            }
        };

        for (int s = 0; s < Tmax; s++) {
            const int slot = s % V5_RING;
            const float *xs_ = reinterpret_cast<const float *>(gring + slot * C::SLOT_B) + jj;
            mbar_wait(&bar_x[slot], (s / V5_RING) & 1);            // this step's input columns have landed
            if (q == 0) V6_TRACE(5);

            // reset gate -> (r * h) operand
            mbar_wait(bar_r, s & 1);
            tc_fence_after();
            if (q == 0) V6_TRACE(6);
            {
                f32x2 t[NP], gr[NP], rh[NP];
                preact(0, xs_ + H, k_sig, t);
                logistic_pk<MATH, NP>(t, gr);
#pragma unroll
                for (int p = 0; p < NP; p++) rh[p] = mul2(gr[p], hs[p]);
                write_operand(b_rh, rh);
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_rh);
            if (q == 0) V6_TRACE(7);

            // update gate (its UMMAs ran while the reset gate was being evaluated).  MATH 5 keeps u = e^-a and 1 + u
            // instead of z: the state update below needs ONE reciprocal for z and tanh together.
            f32x2 gz[NP], gu[NP];
            mbar_wait(bar_z, s & 1);
            tc_fence_after();
            if (q == 0) V6_TRACE(8);
            {
                f32x2 t[NP];
                preact(NM, xs_, k_sig, t);
                if (MATH == 0) {
                    logistic_pk<MATH, NP>(t, gz);
                } else {
#pragma unroll
                    for (int p = 0; p < NP; p++) {
                        float t0, t1;
                        upk2(t[p], t0, t1);
                        gu[p] = pk2(ex2_approx(fminf(t0, 40.0f)), ex2_approx(fminf(t1, 40.0f)));
                        gz[p] = add2(gu[p], splat2(1.0f));                  // A = 1 + u (not z)
                    }
                }
            }

            if (q == 0) V6_TRACE(9);
            // candidate, state update, next step's operand, result row
            mbar_wait(bar_c, s & 1);
            tc_fence_after();
            if (q == 0) V6_TRACE(10);
            {
                f32x2 t[NP];
                preact(2 * NM, xs_ + 2 * H, k_tanh, t);
                float *orow = reinterpret_cast<float *>(gout + (s & 1) * C::OUT_B) + jj;
                if (MATH == 0) {
#pragma unroll
                    for (int p = 0; p < NP; p++) {
                        float t0, t1;
                        upk2(t[p], t0, t1);
                        const f32x2 cand = pk2(tanh_cephes(t0), tanh_cephes(t1));
                        const f32x2 omz = fma2(gz[p], splat2(-1.0f), splat2(1.0f));     // 1 - z
                        // h' = z h + (1 - z) cand, on the pre-scaled state: hs' = z hs + (1 - z) (256 cand)
                        hs[p] = fma2(gz[p], hs[p], mul2(omz, mul2(cand, splat2(OPERAND_SCALE))));
                    }
                } else {
                    // z = 1 / (1 + u), tanh = (1 - v) / (1 + v) with u = e^-a, v = e^-2b:
                    //   h' = h + (1 - z)(tanh - h) = h + u [(1 - v) - h (1 + v)] / [(1 + u)(1 + v)]
                    // -- one refined reciprocal for both gates (5 MUFU per element and step instead of 6), written as
                    // a CORRECTION to h: a unit whose update gate is shut (u below 2^-24 of the bracket) keeps its
                    // state bit for bit, as with z h + (1 - z) c where z rounds to 1; a quotient form would re-round a
                    // persistent state at every step and let it drift.  Exponents are clamped at 2^40 (sigmoid exact to
                    // 1e-12) so that every intermediate, times the 2^8 operand scale, stays finite.
#pragma unroll
                    for (int p = 0; p < NP; p++) {
                        float t0, t1;
                        upk2(t[p], t0, t1);
                        const f32x2 v = pk2(ex2_approx(fminf(t0, 40.0f)), ex2_approx(fminf(t1, 40.0f)));
                        const f32x2 Bn = fma2(v, splat2(-1.0f), splat2(-1.0f));        // -(1 + v)
                        const f32x2 Dn = mul2(gz[p], Bn);                                 // -(1 + u)(1 + v)
                        float d0, d1;
                        upk2(Dn, d0, d1);
                        const f32x2 q0 = pk2(rcp_approx(-d0), rcp_approx(-d1));
                        const f32x2 q = fma2(q0, fma2(Dn, q0, splat2(1.0f)), q0);       // Newton step
                        const f32x2 w256 = mul2(add2(Bn, splat2(2.0f)), splat2(OPERAND_SCALE));   // 256 (1 - v)
                        const f32x2 br = fma2(hs[p], Bn, w256);                           // 256 (1 - v) - hs (1 + v)
                        hs[p] = fma2(mul2(gu[p], br), q, hs[p]);
                    }
                }
#pragma unroll
                for (int p = 0; p < NP; p++) {
                    f32x2 o = mul2(hs[p], splat2(1.0f / OPERAND_SCALE));
                    if (RESID) o = add2(o, pk2(xs_[(2 * p) * XSTR + 3 * H], xs_[(2 * p + 1) * XSTR + 3 * H]));
                    float o0, o1;
                    upk2(o, o0, o1);
                    if (valid) { orow[(2 * p) * H] = o0; orow[(2 * p + 1) * H] = o1; }
                }
                write_operand(b_h, hs);
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_h);
            if (q == 0) V6_TRACE(11);
        }
    }
#undef V6_TRACE
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

// All 512 TMEM columns are allocated by a scan CTA, so a second one on the same SM would stall in tcgen05.alloc:
// the dynamic shared-memory request (v4: 104 KB on top of 26 KB static; v5: >= 116 KB) keeps it off while leaving
// room for the decode / conv CTAs of other batches that share the SM.
constexpr int V4_EXCLUSIVE_SMEM = 104 * 1024;

template <int H, int MATH, int NG>
static int launch_scan_v4(const float *Xin, const float *sW, const float *sW2, const float *resid, float *out,
                          const BatchDims &d, int backward, long long *trace, cudaStream_t s) {
    const int grid = (d.nread + 4 * NG - 1) / (4 * NG);
    gru_scan_v4_kernel<H, MATH, NG><<<grid, (H > 96) ? 512 : 128 * NG, V4_EXCLUSIVE_SMEM, s>>>(Xin, sW, sW2, resid, out, d, backward, trace);
    return 0;
}

template <int H, int MATH, int NG>
static int launch_scan_v5(const float *Xin, const float *sW, const float *sW2, const float *resid, float *out,
                          const BatchDims &d, int backward, cudaStream_t s) {
    const int grid = (d.nread + 8 * NG - 1) / (8 * NG);
    if (resid != nullptr)
        gru_scan_v5_kernel<H, MATH, NG, true><<<grid, 512, ScanV5Cfg<H, NG, true>::SMEM_REQ, s>>>(Xin, sW, sW2, resid, out, d, backward);
    else
        gru_scan_v5_kernel<H, MATH, NG, false><<<grid, 512, ScanV5Cfg<H, NG, false>::SMEM_REQ, s>>>(Xin, sW, sW2, resid, out, d, backward);
    return 0;
}

template <int H, int MATH, int NG, int RPG>
static int launch_scan_v6(const float *Xin, const float *sW, const float *sW2, const float *resid, float *out,
                          const BatchDims &d, int backward, long long *trace, cudaStream_t s) {
    const int grid = (d.nread + RPG * NG - 1) / (RPG * NG);
    if (resid != nullptr)
        gru_scan_v6_kernel<H, MATH, NG, RPG, true><<<grid, ScanV6Cfg<H, NG, RPG, true>::NTHREADS, ScanV6Cfg<H, NG, RPG, true>::SMEM_REQ, s>>>(
            Xin, sW, sW2, resid, out, d, backward, trace);
    else
        gru_scan_v6_kernel<H, MATH, NG, RPG, false><<<grid, ScanV6Cfg<H, NG, RPG, false>::NTHREADS, ScanV6Cfg<H, NG, RPG, false>::SMEM_REQ, s>>>(
            Xin, sW, sW2, resid, out, d, backward, trace);
    return 0;
}

// gen: 4 / 5 force the v4 / v5 kernel (sb2_engine_set_scan_generation, or SCRAPPIE_B200_SCAN_GEN at engine creation);
// 0 = v5 for batches of >= 48 reads, v4 below -- a short step matters more than SM time when one CTA holds the whole
// batch.  SCRAPPIE_B200_SCAN_GROUPS=2|3|4 (read once, thread-safe static initialisation) overrides the v4 groups.
static int scan_env(const char *name) {
    const char *e = getenv(name);
    return e ? atoi(e) : 0;
}

// math: 0 cephes-identical gates, 2 polynomial exp2, 5 SFU ex2 + Newton-refined reciprocal (default)
int launch_gru_scan_tc(const float *Xin, const float *sW, const float *sW2, const float *resid, float *out,
                       const BatchDims &d, int H, int backward, int math, int gen, long long *trace, cudaStream_t s) {
    static const int groups = scan_env("SCRAPPIE_B200_SCAN_GROUPS");
    if ((gen == 0 || gen == 6) && (math == 5 || math == 0)) {
        // v6: eight reads per group once a batch fills at least one and a half CTAs that way, four below
        const bool big = d.nread >= 48;
#define SB2_V6(HH, MM, GG, RR) if (H == HH && math == MM) return launch_scan_v6<HH, MM, GG, RR>(Xin, sW, sW2, resid, out, d, backward, trace, s)
        if (big) { SB2_V6(96, 5, 4, 8); SB2_V6(96, 0, 4, 8); SB2_V6(112, 5, 3, 8); SB2_V6(112, 0, 3, 8); }
        else { SB2_V6(96, 5, 2, 4); SB2_V6(96, 0, 2, 4); SB2_V6(112, 5, 2, 4); SB2_V6(112, 0, 2, 4); }
#undef SB2_V6
    }
    const bool v5 = (gen == 5) || (gen != 4 && d.nread >= 48);
    if (v5 && trace == nullptr) {
#define SB2_V5(HH, MM, GG) if (H == HH && math == MM) return launch_scan_v5<HH, MM, GG>(Xin, sW, sW2, resid, out, d, backward, s)
        SB2_V5(96, 5, 4); SB2_V5(96, 2, 4); SB2_V5(96, 0, 4);
        SB2_V5(112, 5, 3); SB2_V5(112, 2, 3); SB2_V5(112, 0, 3);
#undef SB2_V5
    }
    // v4 reads per CTA: 16 (four groups) once a batch has enough reads (SCRAPPIE_B200_SCAN_GROUPS=2|3|4 overrides)
    const bool four = (H == 96) && (groups == 4 || (groups == 0 && d.nread >= 128));
    const bool three = (H == 112) && (groups == 3 || (groups == 0 && d.nread >= 96));
#define SB2_CASE3(MM) if (three && math == MM) return launch_scan_v4<112, MM, 3>(Xin, sW, sW2, resid, out, d, backward, trace, s)
    SB2_CASE3(5); SB2_CASE3(2); SB2_CASE3(0);
#undef SB2_CASE3
#define SB2_CASE(HH, MM) if (H == HH && math == MM) return launch_scan_v4<HH, MM, 2>(Xin, sW, sW2, resid, out, d, backward, trace, s)
#define SB2_CASE4(MM) if (four && math == MM) return launch_scan_v4<96, MM, 4>(Xin, sW, sW2, resid, out, d, backward, trace, s)
    SB2_CASE4(5); SB2_CASE4(2); SB2_CASE4(0);
    SB2_CASE(96, 0); SB2_CASE(96, 2); SB2_CASE(96, 5);
    SB2_CASE(112, 0); SB2_CASE(112, 2); SB2_CASE(112, 5);
#undef SB2_CASE
#undef SB2_CASE4
    return -1;
}

// Per-device function attributes of this file's kernels (called once per engine, after cudaSetDevice).
template <int H, int MATH>
static bool configure_scan_hm() {
    const cudaFuncAttribute A = cudaFuncAttributeMaxDynamicSharedMemorySize;
    constexpr int NG5 = (H > 96) ? 3 : 4, NG4 = (H > 96) ? 3 : 4;
    bool ok = cudaFuncSetAttribute(gru_scan_v4_kernel<H, MATH, 2>, A, V4_EXCLUSIVE_SMEM) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(gru_scan_v4_kernel<H, MATH, NG4>, A, V4_EXCLUSIVE_SMEM) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(gru_scan_v5_kernel<H, MATH, NG5, false>, A, (int)ScanV5Cfg<H, NG5, false>::SMEM_REQ) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(gru_scan_v5_kernel<H, MATH, NG5, true>, A, (int)ScanV5Cfg<H, NG5, true>::SMEM_REQ) == cudaSuccess;
    return ok;
}

template <int H, int MATH, int NG, int RPG>
static bool configure_scan_v6() {
    const cudaFuncAttribute A = cudaFuncAttributeMaxDynamicSharedMemorySize;
    return cudaFuncSetAttribute(gru_scan_v6_kernel<H, MATH, NG, RPG, false>, A, (int)ScanV6Cfg<H, NG, RPG, false>::SMEM_REQ) == cudaSuccess &&
           cudaFuncSetAttribute(gru_scan_v6_kernel<H, MATH, NG, RPG, true>, A, (int)ScanV6Cfg<H, NG, RPG, true>::SMEM_REQ) == cudaSuccess;
}

int configure_scan_kernels() {
    if (!(configure_scan_v6<96, 5, 4, 8>() && configure_scan_v6<96, 0, 4, 8>() && configure_scan_v6<112, 5, 3, 8>() &&
          configure_scan_v6<112, 0, 3, 8>() && configure_scan_v6<96, 5, 2, 4>() && configure_scan_v6<96, 0, 2, 4>() &&
          configure_scan_v6<112, 5, 2, 4>() && configure_scan_v6<112, 0, 2, 4>()))
        return -1;
    const bool ok = configure_scan_hm<96, 0>() && configure_scan_hm<96, 2>() && configure_scan_hm<96, 5>() &&
                    configure_scan_hm<112, 0>() && configure_scan_hm<112, 2>() && configure_scan_hm<112, 5>();
    return ok ? 0 : -1;
}

}  // namespace sb2
