// Tensor-core (tcgen05 / TMEM) kernels for sm_100a.
//
//   gru_scan_kernel      the recurrent part of a GRU layer (src/layers.c:373-527 in the reference): per time step two
//                        dependent products sW^T h (2H x N) and sW2^T (r*h) (H x N) as UMMA tiles with the gate math
//                        fused between them; the weights stay resident in tensor memory for the whole layer, inputs
//                        arrive through a TMA ring, results leave through TMA stores.
//   tc_selftest_kernel   one UMMA tile product checked against the host (descriptor / layout validation and latency
//                        probe).
//
// Numerics: split-fp16 operands, fp32 accumulation (tc_common.cuh).
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
// GRU scan on tcgen05: gru_forward / gru_backward (src/layers.c:373-527) + residual_inplace (:303-319)
// ---------------------------------------------------------------------------------
// Per time step a GRU layer needs two DEPENDENT products, sW^T h (2H rows) and sW2^T (r * h) (H rows), with the gate
// non-linearities between and after them; the input transform iW^T x + b of all steps is computed beforehand by
// affine_tc_kernel (kernels_gemm.cu).  What the kernel looks like is the result of measuring where a step goes
// (tools/scan_trace.py, profiles/):
//
//   * One CTA = NG independent GROUPS of RPG reads.  A group has its own operands, accumulators, mbarriers, one UMMA
//     issuer warp and NQ gate warps; the groups' dependency chains interleave on the same schedulers and the same
//     tensor pipe, so one group's gate math fills the other groups' UMMA and hand-over latencies.
//   * The recurrent weights (r, z, c tiles, fp16 hi and lo: 6 tiles of H / 2 columns) live in TENSOR MEMORY for the
//     whole layer and are the A operand (tcgen05.mma with A from TMEM): ~9 cycles per M128 N16 K16 instruction
//     against ~38 when A is fetched from shared memory.
//   * The B operand holds the state of the group's reads, split into fp16 hi and lo, SIDE BY SIDE IN N: rows 0..7 are
//     the hi halves of (up to) 8 reads, rows 8..15 the lo halves, so a product is two passes (A = W_lo, A = W_hi)
//     instead of three, and accumulator columns i and 8 + i add up to read i's result.  RPG = 8 uses the
//     instruction's minimum N = 16 fully (batches of >= 48 reads: 32 reads per CTA at H = 96 with NG = 4,
//     288 + 4 x 48 = 480 TMEM columns; 24 at H = 112 with NG = 3); small batches take RPG = 4, NG = 2 (shorter steps).
//   * The reset gate is issued and committed first; z's UMMAs run while the gate warps turn r into the (r * h)
//     operand, and the z exponentials are evaluated while the candidate's UMMAs run.
//   * Gate math: TMEM lane = hidden unit, so a thread owns one unit of RPG reads.  fp32 arithmetic runs on PAIRS of
//     reads with the packed instructions of sm_100 (fma / add / mul .f32x2 -> FFMA2 / FADD2 / FMUL2); the state is
//     kept pre-scaled by 2^8 (the operand scale).  ~370 instructions per step and thread for 8 reads, where a scalar
//     formulation had 579.
//   * Inputs: affine_tc_kernel writes Xin in the scan's own order -- per group, step-major, the group's reads next to
//     each other ([group][step][read][3H], backward layers in reverse time) -- so ONE cp.async.bulk (TMA, SASS
//     UBLKCP) per group and step brings all the group's input columns into a 3-slot shared-memory ring two steps
//     ahead of their use; completion is counted in bytes on the slot's mbarrier.  (Per-read copies from a read-major
//     Xin cost the issuer ~600 cycles per step -- each lane's copy is issued one after the other -- and sat on the
//     critical path.)  Callers that keep Xin read-major (raw_r94, the fp32 debug GEMM) get per-read copies.
//   * Results: the gate warps put a step's output row into a shared-memory staging slot, the issuer stores it with
//     cp.async.bulk after the candidate's UMMAs are on their way (its idle time); the residual input of rnnrf is
//     read by the gate threads themselves (coalesced, off the critical path).
//   * All issuers sit on warps with warp % 4 == 3 where H <= 96 leaves scheduler 3 without gate math.
//
// Arithmetic: split fp16 operands, fp32 accumulation (tc_common.cuh); MATH 5 = ex2.approx + Newton-refined rcp.approx
// (default), MATH 0 = the reference's cephes polynomials element by element (SCRAPPIE_B200_SCAN=cephes).  The same
// code serves RPG = 4 and 8, so a read's result does not depend on the batch it travels in.
// Input slots: the copy for step s + RING - 1 is issued while step s runs.  Eight-read groups step every ~1.4 us, so two
// steps of lead cover HBM latency under load; four-read groups (small batches, long reads) step every ~0.7 us and get
// five steps of lead.
template <int RPG> struct ScanRing { static constexpr int N = (RPG == 8) ? 3 : 6; };

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
            y[p] = pk2(logistic_cephes(t0), logistic_cephes(t1));
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

// how many reads share a group (= a UMMA B operand) for a batch of `nread` reads: eight once the batch fills at least one
// and a half CTAs that way, four below (a short step matters more than SM time when one CTA holds the whole batch)
int scan_reads_per_group(int nread) { return nread >= 48 ? 8 : 4; }

template <int H, int NG, int RPG, bool RESID>
struct ScanCfg {
    static constexpr int NM = 16;
    static constexpr int NQ = (H + 31) / 32;            // TMEM lane quarters that own hidden units
    // H <= 96: four warps per group (gate warps 4g .. 4g+2, issuer 4g+3).  H = 112: gate warps 0 .. 4 NG - 1, one
    // idle warp, issuers 4 NG + 1 + g (never on scheduler 0)
    static constexpr int NTHREADS = (NQ < 4) ? 128 * NG : 32 * (5 * NG + 1);
    static constexpr uint32_t LBO_B = 16u * NM + 16u, SBO_B = 128u;
    static constexpr uint32_t TILE_B = (H / 8) * LBO_B;
    static constexpr uint32_t XCOL_B = 3 * H * 4;                                  // input bytes per read and step
    static constexpr uint32_t SLOT_B = RPG * XCOL_B;                               // one group, one step
    static constexpr uint32_t OUT_B = RPG * H * 4;                                 // one group, one step of results
    static constexpr uint32_t OFF_RING = (NG * 2 * TILE_B + 127) / 128 * 128;
    static constexpr int SCAN_RING = ScanRing<RPG>::N;
    static constexpr uint32_t OFF_OUT = OFF_RING + NG * SCAN_RING * SLOT_B;
    static constexpr uint32_t OFF_BAR = OFF_OUT + NG * 2 * OUT_B;
    static constexpr uint32_t NBAR = NG * (5 + SCAN_RING);
    static constexpr uint32_t OFF_META = OFF_BAR + NBAR * 8 + 16;                  // per group: first column [8], length [8]
    static constexpr uint32_t SMEM = OFF_META + NG * 16 * 4;
    // all 512 TMEM columns are allocated: a second scan CTA on the SM would stall in tcgen05.alloc, so the request
    // is at least half of the SM's shared memory -- which still leaves ~100 KB for the decode / conv CTAs of other
    // batches that share the SM
    static constexpr uint32_t SMEM_REQ = SMEM > 116 * 1024 ? SMEM : 116 * 1024;
    static_assert(SLOT_B % 16 == 0 && XCOL_B % 16 == 0 && OUT_B % 16 == 0 && (H * 4) % 16 == 0, "bulk copy alignment");
    static_assert(SMEM <= 200 * 1024, "shared memory budget");
};

// XIL: Xin is in the scan's own order (one bulk copy per group and step); xgrp[g] = first row of group g.
template <int H, int MATH, int NG, int RPG, bool RESID, bool XIL>
__global__ void __launch_bounds__(ScanCfg<H, NG, RPG, RESID>::NTHREADS, 1)
gru_scan_kernel(const float *__restrict__ Xin, const long long *__restrict__ xgrp, const float *__restrict__ sW,
                const float *__restrict__ sW2, const float *__restrict__ resid, float *__restrict__ out, BatchDims d,
                int backward, long long *__restrict__ trace) {
    using C = ScanCfg<H, NG, RPG, RESID>;
    // diagnostic (SCRAPPIE_B200_TRACE=1): clock64() of CTA 0 at the hand-over points of steps 100..103, per group:
    // trace[(grp * 4 + step - 100) * 16 + slot]; slots 0-4 issuer, 5-11 gate warp of lane quarter 0
#define SCAN_TRACE(slot) do { if (trace != nullptr && blockIdx.x == 0 && lane == 0 && s >= 100 && s < 104) trace[((grp * 4) + (s - 100)) * 16 + (slot)] = clock64(); } while (0)
    constexpr int NM = C::NM, NP = RPG / 2, NQ = C::NQ, SCAN_RING = C::SCAN_RING;
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

    // operands and input ring start as zeros (rows of reads that do not exist stay zero)
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
            uint64_t *gb = bars + g * (5 + SCAN_RING);
            mbar_init(&gb[0], 1);                       // r committed
            mbar_init(&gb[1], 1);                       // z committed
            mbar_init(&gb[2], 1);                       // c committed
            mbar_init(&gb[3], NQ);                      // r*h operand written
            mbar_init(&gb[4], NQ);                      // h operand + result row written
            for (int k = 0; k < SCAN_RING; k++) mbar_init(&gb[5 + k], 1);   // input slot k filled (transaction bytes)
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

    const bool is_issuer = (NQ < 4) ? ((warp & 3) == 3 && warp < 4 * NG) : (warp > 4 * NG && warp <= 5 * NG);
    const bool is_gate = (warp < 4 * NG) && ((warp & 3) < NQ);
    const int grp = (NQ < 4) ? (warp >> 2) : (is_issuer ? (warp - 4 * NG - 1) : (warp >> 2));
    const int gsel = (is_issuer || is_gate) ? grp : 0;
    uint8_t *b_h = b_ops + gsel * 2 * TILE_B, *b_rh = b_h + TILE_B;
    uint64_t *gb = bars + gsel * (5 + SCAN_RING);
    uint64_t *bar_r = &gb[0], *bar_z = &gb[1], *bar_c = &gb[2], *bar_rh = &gb[3], *bar_h = &gb[4], *bar_x = &gb[5];
    uint8_t *gring = ring + (size_t)gsel * SCAN_RING * C::SLOT_B;
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
        float *odst = out + (size_t)mycol * H;
        // (groups past the end of the batch exist in the last CTA: Tmax = 0, nothing is copied, no table entry is read)
        const float *xsrc = XIL ? (Xin + (size_t)(Tmax > 0 ? xgrp[blockIdx.x * NG + grp] : 0) * (3 * H)) : (Xin + (size_t)mycol * (3 * H));
        auto fill = [&](int st) {                       // request the inputs of step st (all lanes call it)
            if (st < Tmax) {
                const int slot = st % SCAN_RING;
                if (XIL) {
                    // one copy: the group's RPG input columns of a step are contiguous in the scan-ordered Xin
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&bar_x[slot], C::SLOT_B);
                        bulk_g2s(gring + slot * C::SLOT_B, xsrc + (size_t)st * (RPG * 3 * H), C::SLOT_B, &bar_x[slot]);
                    }
                } else {
                    const bool mine = st < myT;         // lanes >= RPG have myT = 0
                    const unsigned vm = __ballot_sync(0xffffffffu, mine);
                    if (lane == 0) mbar_arrive_expect_tx(&bar_x[slot], (uint32_t)__popc(vm) * C::XCOL_B);
                    __syncwarp();
                    if (mine) bulk_g2s(gring + slot * C::SLOT_B + lane * C::XCOL_B, xsrc + (ptrdiff_t)st * dir * (3 * H), C::XCOL_B, &bar_x[slot]);
                }
            }
            __syncwarp();
        };
        auto store = [&](int st) {                      // write the results of step st (staged by the gate warps)
            if (st < myT) bulk_s2g(odst + (ptrdiff_t)st * dir * H, gout + (st & 1) * C::OUT_B + lane * (H * 4), H * 4);
            bulk_commit();
            __syncwarp();
        };
#pragma unroll 1
        for (int st = 0; st < SCAN_RING - 1; st++) fill(st);
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
            SCAN_TRACE(0);
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
            SCAN_TRACE(1);
            // every gate warp has finished step s - 1 (bar_h), so the input slot that step used is free again
            fill(s + SCAN_RING - 1);
            SCAN_TRACE(2);
            mbar_wait(bar_rh, s & 1);
            tc_fence_after();
            SCAN_TRACE(3);
            // the staging row the gate warps fill after THIS commit was last read by the store issued a step ago
            bulk_wait_read0();
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
            SCAN_TRACE(4);
            // idle until the gate warps finish the step: the results of step s - 1 go out now
            if (s > 0) store(s - 1);
        }
        if (Tmax > 0) {
            mbar_wait(bar_h, (Tmax - 1) & 1);
            bulk_wait_read0();
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
        constexpr int XSTR = 3 * H;                     // floats between consecutive reads of an input slot
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
            const int slot = s % SCAN_RING;
            const float *xs_ = reinterpret_cast<const float *>(gring + slot * C::SLOT_B) + jj;
            // residual input (rnnrf): this thread's element of each read's column, needed only at the end of the step
            f32x2 rs[NP];
            if (RESID) {
#pragma unroll
                for (int p = 0; p < NP; p++) {
                    float r0_ = 0.0f, r1_ = 0.0f;
                    if (s < gmeta[8 + 2 * p]) r0_ = __ldg(resid + (size_t)(gmeta[2 * p] + s * dir) * H + jj);
                    if (s < gmeta[8 + 2 * p + 1]) r1_ = __ldg(resid + (size_t)(gmeta[2 * p + 1] + s * dir) * H + jj);
                    rs[p] = pk2(r0_, r1_);
                }
            }
            mbar_wait(&bar_x[slot], (s / SCAN_RING) & 1);          // this step's input columns have landed
            if (q == 0) SCAN_TRACE(5);

            // reset gate -> (r * h) operand
            mbar_wait(bar_r, s & 1);
            tc_fence_after();
            if (q == 0) SCAN_TRACE(6);
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
            if (q == 0) SCAN_TRACE(7);

            // update gate (its UMMAs ran while the reset gate was being evaluated)
            f32x2 gz[NP];
            mbar_wait(bar_z, s & 1);
            tc_fence_after();
            if (q == 0) SCAN_TRACE(8);
            {
                f32x2 t[NP];
                preact(NM, xs_, k_sig, t);
                logistic_pk<MATH, NP>(t, gz);
            }

            if (q == 0) SCAN_TRACE(9);
            // candidate, state update, next step's operand, result row
            mbar_wait(bar_c, s & 1);
            tc_fence_after();
            if (q == 0) SCAN_TRACE(10);
            {
                f32x2 t[NP], cand[NP];
                preact(2 * NM, xs_ + 2 * H, k_tanh, t);
                float *orow = reinterpret_cast<float *>(gout + (s & 1) * C::OUT_B) + jj;
                if (MATH == 0) {
#pragma unroll
                    for (int p = 0; p < NP; p++) {
                        float t0, t1;
                        upk2(t[p], t0, t1);
                        cand[p] = pk2(tanh_cephes(t0), tanh_cephes(t1));
                    }
                } else {
                    f32x2 sg[NP];
                    logistic_pk<MATH, NP>(t, sg);       // tanh(x) = 2 sigmoid(2 x) - 1 (the doubling is in k_tanh)
#pragma unroll
                    for (int p = 0; p < NP; p++) cand[p] = fma2(sg[p], splat2(2.0f), splat2(-1.0f));
                }
                // h' = z h + (1 - z) cand on the pre-scaled state: hs' = z hs + (1 - z)(256 cand).  (Sharing one
                // reciprocal between z and tanh -- h' = h + u [(1 - v) - h (1 + v)] / [(1 + u)(1 + v)] -- saves a
                // MUFU per element but was measured 3-4 x less accurate where a state is replaced (z ~ 0): the
                // quotient's relative error multiplies |cand - h| instead of each term's own rounding.  Not used.)
#pragma unroll
                for (int p = 0; p < NP; p++) {
                    const f32x2 omz = fma2(gz[p], splat2(-1.0f), splat2(1.0f));         // 1 - z
                    hs[p] = fma2(gz[p], hs[p], mul2(omz, mul2(cand[p], splat2(OPERAND_SCALE))));
                }
#pragma unroll
                for (int p = 0; p < NP; p++) {
                    f32x2 o = mul2(hs[p], splat2(1.0f / OPERAND_SCALE));
                    if (RESID) o = add2(o, rs[p]);
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
            if (q == 0) SCAN_TRACE(11);
        }
    }
#undef SCAN_TRACE
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

// src_col tables for affine_tc_kernel: which input column (read r, time t) each row of the scan-ordered Xin is made
// from, for forward layers (step = t) and backward layers (step = T - 1 - t).  One CTA per read; rows of a ragged group
// that no read reaches keep the -1 the tables were filled with.
__global__ void scan_rows_kernel(BatchDims d, const long long *__restrict__ xgrp, int rpg, int *__restrict__ src_f,
                                 int *__restrict__ src_b) {
    const int r = blockIdx.x;
    const int T = d.nblock[r], col = d.col_off[r];
    const long long base = xgrp[r / rpg] + (r % rpg);
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        src_f[base + (long long)t * rpg] = col + t;
        src_b[base + (long long)(T - 1 - t) * rpg] = col + t;
    }
}

int launch_scan_rows(const BatchDims &d, const long long *xgrp, int rpg, size_t nrow, int *src_f, int *src_b, cudaStream_t s) {
    if (cudaMemsetAsync(src_f, 0xFF, nrow * sizeof(int), s) != cudaSuccess || cudaMemsetAsync(src_b, 0xFF, nrow * sizeof(int), s) != cudaSuccess)
        return -1;
    if (d.nread > 0) scan_rows_kernel<<<d.nread, 256, 0, s>>>(d, xgrp, rpg, src_f, src_b);
    return 0;
}

// 0 = launches inherit their stream's priority; otherwise the device's greatest priority (a negative number)
int scan_launch_priority() {
    const char *e = getenv("SCRAPPIE_B200_SCAN_PRIO");
    if (e == nullptr || atoi(e) == 0) return 0;
    int least = 0, greatest = 0;
    if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) return 0;
    return greatest;
}

template <int H, int MATH, int NG, int RPG, bool RESID>
static int launch_scan_cfg(const float *Xin, const long long *xgrp, const float *sW, const float *sW2, const float *resid,
                           float *out, const BatchDims &d, int backward, long long *trace, cudaStream_t s) {
    using C = ScanCfg<H, NG, RPG, RESID>;
    const int grid = (d.nread + RPG * NG - 1) / (RPG * NG);
    // SCRAPPIE_B200_SCAN_PRIO=1 (experiment, read once): the scan's CTAs are dispatched ahead of other batches' pending
    // CTAs -- the scans are the long pole of every batch's chain and need a completely free SM each
    static const int prio = scan_launch_priority();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(C::NTHREADS); cfg.dynamicSmemBytes = C::SMEM_REQ; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    if (prio != 0) {
        attr[0].id = cudaLaunchAttributePriority;
        attr[0].val.priority = prio;
        cfg.attrs = attr; cfg.numAttrs = 1;
    }
    cudaError_t e;
    if (xgrp != nullptr)
        e = cudaLaunchKernelEx(&cfg, gru_scan_kernel<H, MATH, NG, RPG, RESID, true>, Xin, xgrp, sW, sW2, resid, out, d, backward, trace);
    else
        e = cudaLaunchKernelEx(&cfg, gru_scan_kernel<H, MATH, NG, RPG, RESID, false>, Xin, xgrp, sW, sW2, resid, out, d, backward, trace);
    return e == cudaSuccess ? 0 : -1;
}

// math: 0 cephes-identical gates, 5 SFU ex2 + Newton-refined reciprocal (default).  xgrp: first Xin row of every read
// group when Xin is in scan order (see affine_tc_kernel's src_col), nullptr when Xin is read-major.  Instantiated for
// the shapes the models have: H = 96 without and H = 112 with the residual input.
int launch_gru_scan_tc(const float *Xin, const long long *xgrp, const float *sW, const float *sW2, const float *resid,
                       float *out, const BatchDims &d, int H, int backward, int math, long long *trace, cudaStream_t s) {
    const bool big = scan_reads_per_group(d.nread) == 8;
#define SB2_SCAN(HH, MM, GG, RR, RES) \
    if (H == HH && math == MM && big == (RR == 8) && (resid != nullptr) == RES) \
        return launch_scan_cfg<HH, MM, GG, RR, RES>(Xin, xgrp, sW, sW2, resid, out, d, backward, trace, s)
    SB2_SCAN(96, 5, 4, 8, false); SB2_SCAN(96, 5, 2, 4, false); SB2_SCAN(96, 0, 4, 8, false); SB2_SCAN(96, 0, 2, 4, false);
    SB2_SCAN(112, 5, 3, 8, true); SB2_SCAN(112, 5, 2, 4, true); SB2_SCAN(112, 0, 3, 8, true); SB2_SCAN(112, 0, 2, 4, true);
#undef SB2_SCAN
    return -1;
}

// Per-device function attributes of this file's kernels (called once per engine, after cudaSetDevice).
template <int H, int MATH, int NG, int RPG, bool RESID>
static bool configure_scan_cfg() {
    const cudaFuncAttribute A = cudaFuncAttributeMaxDynamicSharedMemorySize;
    const int smem = (int)ScanCfg<H, NG, RPG, RESID>::SMEM_REQ;
    return cudaFuncSetAttribute(gru_scan_kernel<H, MATH, NG, RPG, RESID, true>, A, smem) == cudaSuccess &&
           cudaFuncSetAttribute(gru_scan_kernel<H, MATH, NG, RPG, RESID, false>, A, smem) == cudaSuccess;
}

int configure_scan_kernels() {
    const bool ok = configure_scan_cfg<96, 5, 4, 8, false>() && configure_scan_cfg<96, 5, 2, 4, false>() &&
                    configure_scan_cfg<96, 0, 4, 8, false>() && configure_scan_cfg<96, 0, 2, 4, false>() &&
                    configure_scan_cfg<112, 5, 3, 8, true>() && configure_scan_cfg<112, 5, 2, 4, true>() &&
                    configure_scan_cfg<112, 0, 3, 8, true>() && configure_scan_cfg<112, 0, 2, 4, true>();
    return ok ? 0 : -1;
}

}  // namespace sb2
