// Time-batched affine maps on the 5th-generation tensor cores (tcgen05 / TMEM).
//
//   affine_tc_kernel   C[col][m] = b[m] + sum_k W[m][k] X[col][k]       (feedforward_linear ->
//                      affine_map, src/layers.c:248-252, src/scrappie_matrix.c:323-351): the input
//                      transform of every GRU layer for ALL time steps of ALL reads of a batch.
//
// Roles inside a CTA (persistent, one CTA per SM, 416 threads):
//   warps 5-12 producers  load a chunk of NT activation columns (fp32, coalesced), split each value
//                         into fp16 hi + lo and store it as the canonical K-major UMMA B operand
//   warp  4    issuer     one elected lane issues tcgen05.mma: A = weight tile (shared memory,
//                         resident for the whole kernel, fetched once by a TMA bulk copy),
//                         B = activation chunk, D = fp32 accumulator in TMEM (double buffered)
//   warps 0-3  epilogue   tcgen05.ld the accumulator (TMEM lane = output unit), add the bias and
//                         store: for a fixed column the 32 lanes of a warp write 128 contiguous bytes
//
// Numerics: split-fp16, three passes (lo*hi, hi*lo, hi*hi), fp32 accumulation -- see tc_common.cuh.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "device_math.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace sb2 {

using namespace tc;

// ---------------------------------------------------------------------------------
// host: weight images
// ---------------------------------------------------------------------------------
// `ntile` tiles of `rows` output units each; per tile a hi and a lo fp16 matrix of rows x K in the
// canonical K-major layout (core matrix = 8 rows x 16 bytes; LBO 128, SBO (K/8)*128).  The UMMA
// reads 128 rows per tile: with rows < 128 it runs into the next tile / the zeroed slack at the end,
// and those accumulator lanes are never read back.
size_t gemm_image_bytes(int ntile, int rows, int K) {
    const size_t tile = (size_t)rows * K * 2;
    const size_t slack = (size_t)((128 - rows) / 8) * (size_t)(K / 8) * 128;
    return 2 * (size_t)ntile * tile + slack;
}

void build_gemm_image(const float *W, int ldw, int M, int K, int rows, int ntile, uint8_t *img) {
    const uint32_t lbo = 128, sbo = (uint32_t)(K / 8) * 128;
    const size_t tile = (size_t)rows * K * 2;
    memset(img, 0, gemm_image_bytes(ntile, rows, K));
    for (int t = 0; t < ntile; t++) {
        uint8_t *hi_t = img + (size_t)(2 * t) * tile, *lo_t = hi_t + tile;
        for (int r = 0; r < rows; r++) {
            const int m = t * rows + r;
            if (m >= M) break;
            for (int k = 0; k < K; k++) {
                const float xs = W[(size_t)m * ldw + k] * OPERAND_SCALE;
                const __half hi = __float2half_rn(xs);
                const __half lo = __float2half_rn(xs - __half2float(hi));
                const uint32_t off = canon_off((uint32_t)r, (uint32_t)k, lbo, sbo);
                *reinterpret_cast<__half *>(hi_t + off) = hi;
                *reinterpret_cast<__half *>(lo_t + off) = lo;
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// device
// ---------------------------------------------------------------------------------
// Columns per chunk of the H = 96 input transform.  64 (159 KB of shared memory) measures the same as 128 (216 KB) and
// lets the CTA share an SM with a decode CTA of another batch instead of waiting for an empty one.
#ifndef SB2_AFFINE_NT
#define SB2_AFFINE_NT 64
#endif
template <int K, int ROWS, int NTILE, int NT>
struct GemmCfg {
    static constexpr uint32_t LBO_A = 128, SBO_A = (K / 8) * 128;
    static constexpr uint32_t TILE_A = ROWS * K * 2;
    static constexpr uint32_t SLACK = ((128 - ROWS) / 8) * SBO_A;
    static constexpr uint32_t WBYTES = 2 * NTILE * TILE_A + SLACK;
    static constexpr uint32_t LBO_B = 16 * NT + 16, SBO_B = 128;
    static constexpr uint32_t TILE_B = (K / 8) * LBO_B;
    static constexpr uint32_t STAGE_B = 2 * TILE_B;            // hi, lo
    static constexpr uint32_t SMEM = WBYTES + 2 * STAGE_B + 128 + 32;   // barriers + TMEM slot, then 4 scales + 3 maxima
    static constexpr int NKS = K / 16;
    static constexpr uint32_t TCOLS = (2 * NT <= 32) ? 32 : (2 * NT <= 64 ? 64 : (2 * NT <= 128 ? 128 : (2 * NT <= 256 ? 256 : 512)));
    static_assert(K % 16 == 0 && ROWS % 8 == 0 && ROWS <= 128 && NT % 16 == 0 && NT <= 256, "tile shape");
    static_assert(WBYTES % 16 == 0 && STAGE_B % 16 == 0, "alignment");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// split 8 consecutive K values into the hi / lo 16-byte rows of a core matrix
__device__ __forceinline__ void split8(const float4 &a, const float4 &b, float scale, uint4 &hi, uint4 &lo) {
    const float x[8] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale, b.x * scale, b.y * scale, b.z * scale, b.w * scale};
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const __half h0 = __float2half_rn(x[2 * i]), h1 = __float2half_rn(x[2 * i + 1]);
        const __half l0 = __float2half_rn(x[2 * i] - __half2float(h0)), l1 = __float2half_rn(x[2 * i + 1] - __half2float(h1));
        h[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        l[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int K, int ROWS, int NTILE, int NT, int LDC>
__global__ void __launch_bounds__(416, 1)
affine_tc_kernel(const float *__restrict__ X, int ncol, const uint8_t *__restrict__ wimg, const float *__restrict__ bias,
                 int M, float *__restrict__ C, const int *__restrict__ src_col) {
    using G = GemmCfg<K, ROWS, NTILE, NT>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *w_img = smem;
    uint8_t *b_ring = smem + G::WBYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(b_ring + 2 * G::STAGE_B);
    uint64_t *full = bars, *empty = bars + 2, *accf = bars + 4, *acce = bars + 6, *wbar = bars + 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 9);
    // Range safety of the split-fp16 operands: activations are scaled by 2^8 before the fp16 split, so |x| >= 2^7.99
    // would overflow to inf where the reference's fp32 GEMM stays finite (conv + ELU output and rnnrf's residual
    // stream are unbounded).  The producers take the chunk's max |x|; chunks below 128 use the usual 2^8 (bit-identical
    // to a build without this guard), others the power of two that puts the maximum in [2^14, 2^15).  The epilogue
    // undoes it.  cscale[it % 4] = 2^-8 (weights) / operand scale of chunk `it`; cmax[it % 3] = bits of max |x|.
    float *cscale = reinterpret_cast<float *>(b_ring + 2 * G::STAGE_B + 128);
    uint32_t *cmax = reinterpret_cast<uint32_t *>(cscale + 4);

    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int nchunk = (ncol + NT - 1) / NT;

    if (tid == 0) {
        cmax[0] = 0; cmax[1] = 0; cmax[2] = 0;
        mbar_init(&full[0], 8); mbar_init(&full[1], 8);
        mbar_init(&empty[0], 1); mbar_init(&empty[1], 1);
        mbar_init(&accf[0], 1); mbar_init(&accf[1], 1);
        mbar_init(&acce[0], 4); mbar_init(&acce[1], 4);
        mbar_init(wbar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, G::TCOLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 4) {
        // ---- weights: one TMA bulk copy stream, then the UMMA issue loop ----------------
        if (lane == 0) {
            mbar_arrive_expect_tx(wbar, G::WBYTES);
            constexpr uint32_t PIECE = 32768;
            for (uint32_t off = 0; off < G::WBYTES; off += PIECE)
                bulk_g2s(w_img + off, wimg + off, (G::WBYTES - off < PIECE) ? (G::WBYTES - off) : PIECE, wbar);
        }
        __syncwarp();
        mbar_wait(wbar, 0);
        const uint32_t idesc = umma_idesc_f16(128, NT);
        const uint64_t dW = umma_desc(smem_u32(w_img), G::LBO_A, G::SBO_A);
        constexpr uint64_t TA = G::TILE_A >> 4, KA = (2 * G::LBO_A) >> 4, KB = (2 * G::LBO_B) >> 4, TB = G::TILE_B >> 4;
        uint32_t it = 0, acc_it = 0;
        for (int c = blockIdx.x; c < nchunk; c += gridDim.x, it++) {
            const uint32_t s = it & 1;
            mbar_wait(&full[s], (it >> 1) & 1);
            const uint64_t dB = umma_desc(smem_u32(b_ring + s * G::STAGE_B), G::LBO_B, G::SBO_B);   // hi; lo at + TB
#pragma unroll 1
            for (int g = 0; g < NTILE; g++, acc_it++) {
                const uint32_t a = acc_it & 1;
                mbar_wait(&acce[a], ((acc_it >> 1) & 1) ^ 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t dcol = tmem + a * NT;
                    const uint64_t w_hi = dW + (uint64_t)(2 * g) * TA, w_lo = w_hi + TA;
#pragma unroll
                    for (int ks = 0; ks < G::NKS; ks++) umma_f16(dcol, w_lo + ks * KA, dB + ks * KB, idesc, ks > 0);
#pragma unroll
                    for (int ks = 0; ks < G::NKS; ks++) umma_f16(dcol, w_hi + ks * KA, dB + TB + ks * KB, idesc, 1);
#pragma unroll
                    for (int ks = 0; ks < G::NKS; ks++) umma_f16(dcol, w_hi + ks * KA, dB + ks * KB, idesc, 1);
                    umma_commit(&accf[a]);
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(&empty[s]);
            __syncwarp();
        }
    } else if (warp >= 5) {
        // ---- producers: fp32 activations -> split fp16 canonical B operand ---------------
        // eight producer warps, and every load of a thread's share of a chunk is issued before the first value is
        // converted (the kernel was bound by the exposed latency of these loads, profiles/r24)
        const int pt = tid - 160;                       // 0..255
        constexpr int K8 = K / 8;
        constexpr int UNITS = NT * K8;
        constexpr int NPROD = 256;
        constexpr int PER = (UNITS + NPROD - 1) / NPROD;
        // Source column of every unit this thread converts.  src_col == nullptr: output row = input column (read-major
        // Xin).  Otherwise the kernel runs over the rows of the SCAN-ORDERED Xin ([group][step][read], kernels_tc.cu) and
        // src_col[row] names the input column that row is made from (-1: a ragged group's unused row): the output
        // stays one contiguous block per chunk, the input becomes a gather of 16-step pieces of 8 reads.  The row
        // numbers of a chunk are fetched while the previous chunk is converted, so no load waits for another.
        // TWO chunks of loads are in flight per thread (register sets A and B, the loop is unrolled by two): while chunk
        // c is converted, the loads of chunk c + grid are already out and the source columns of chunk c + 2 grid are
        // being fetched.  A CTA then pulls twice the bytes per unit of time, so the same HBM bandwidth needs fewer SMs --
        // which is what counts when the scans of other batches hold part of the GPU.
        struct Chunk { float4 va[PER], vb[PER]; };
        auto fetch_cols = [&](int c, int (&scol)[PER]) {
#pragma unroll
            for (int i = 0; i < PER; i++) {
                const int u = pt + i * NPROD;
                const int row = c * NT + u / K8;
                scol[i] = -1;
                if (u < UNITS && c < nchunk && row < ncol) scol[i] = (src_col != nullptr) ? __ldg(src_col + row) : row;
            }
        };
        auto load_chunk = [&](const int (&scol)[PER], Chunk &ch) {
#pragma unroll
            for (int i = 0; i < PER; i++) {
                const int u = pt + i * NPROD;
                ch.va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                ch.vb[i] = ch.va[i];
                if (scol[i] >= 0) {
                    const float *src = X + (size_t)scol[i] * K + (u % K8) * 8;
                    ch.va[i] = __ldg(reinterpret_cast<const float4 *>(src));
                    ch.vb[i] = __ldg(reinterpret_cast<const float4 *>(src + 4));
                }
            }
        };
        auto convert_chunk = [&](const Chunk &ch, uint32_t it) {
            const uint32_t s = it & 1;
            // chunk-wide max |x| (integer compare of the sign-stripped bits; NaN / inf sort above every finite value)
            uint32_t mx = 0;
#pragma unroll
            for (int i = 0; i < PER; i++) {
                const float4 a = ch.va[i], b = ch.vb[i];
                const float m = fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))),
                                      fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
                mx = max(mx, __float_as_uint(m));
                // a NaN is dropped by fmaxf: catch it through the exponent test on each element's bits instead
                mx = max(mx, max(max(__float_as_uint(a.x), __float_as_uint(a.y)), max(__float_as_uint(a.z), __float_as_uint(a.w))) & 0x7fffffffu);
                mx = max(mx, max(max(__float_as_uint(b.x), __float_as_uint(b.y)), max(__float_as_uint(b.z), __float_as_uint(b.w))) & 0x7fffffffu);
            }
            mx = __reduce_max_sync(0xffffffffu, mx);
            const uint32_t ms = it % 3;
            if (lane == 0 && mx >= 0x43000000u) atomicMax(&cmax[ms], mx);      // only chunks with |x| >= 128 touch the slot
            if (pt == 0) cmax[(it + 1) % 3] = 0;                               // last read two chunks ago (see above)
            named_bar_sync(1, NPROD);
            const uint32_t cm = cmax[ms];
            float opscale = OPERAND_SCALE;
            if (cm >= 0x43000000u && cm < 0x7f800000u)                          // finite, >= 128: max * scale in [2^14, 2^15)
                opscale = __uint_as_float((uint32_t)(127 + 14 + 127 - (int)(cm >> 23)) << 23);
            if (pt == 0) cscale[it & 3] = (1.0f / OPERAND_SCALE) / opscale;
            mbar_wait(&empty[s], ((it >> 1) & 1) ^ 1);
            uint8_t *b_hi = b_ring + s * G::STAGE_B, *b_lo = b_hi + G::TILE_B;
#pragma unroll
            for (int i = 0; i < PER; i++) {
                const int u = pt + i * NPROD;
                if (u < UNITS) {
                    const int n = u / K8, k8 = u % K8;
                    uint4 hi, lo;
                    split8(ch.va[i], ch.vb[i], opscale, hi, lo);
                    const uint32_t off = (uint32_t)(n >> 3) * G::SBO_B + (uint32_t)k8 * G::LBO_B + (uint32_t)(n & 7) * 16;
                    *reinterpret_cast<uint4 *>(b_hi + off) = hi;
                    *reinterpret_cast<uint4 *>(b_lo + off) = lo;
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
        };
        const int G_ = (int)gridDim.x;
        Chunk A, B;
        int colsA[PER], colsB[PER];
        uint32_t it = 0;
        int c = blockIdx.x;
        fetch_cols(c, colsA);
        fetch_cols(c + G_, colsB);
        load_chunk(colsA, A);
        while (c < nchunk) {
            load_chunk(colsB, B);                       // chunk c + G (all zeros past the end)
            fetch_cols(c + 2 * G_, colsA);
            convert_chunk(A, it++);
            c += G_;
            if (c >= nchunk) break;
            load_chunk(colsA, A);                       // chunk c + G
            fetch_cols(c + 2 * G_, colsB);
            convert_chunk(B, it++);
            c += G_;
        }
    } else {
        // ---- epilogue: accumulator + bias -> global ---------------------------------------
        const int m = tid;                              // TMEM lane = row inside the tile
        const bool warp_valid = (warp * 32) < ROWS;
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        float bg[NTILE];
        bool ok[NTILE];
#pragma unroll
        for (int g = 0; g < NTILE; g++) {
            ok[g] = (m < ROWS) && (g * ROWS + m < M);
            bg[g] = ok[g] ? bias[g * ROWS + m] : 0.0f;
        }
        uint32_t acc_it = 0, it = 0;
        for (int c = blockIdx.x; c < nchunk; c += gridDim.x, it++) {
            const int col0 = c * NT;
#pragma unroll
            for (int g = 0; g < NTILE; g++, acc_it++) {
                const uint32_t a = acc_it & 1;
                mbar_wait(&accf[a], (acc_it >> 1) & 1);
                tc_fence_after();
                // written by the producers before they released this chunk's operand; 2^-16 unless the chunk held |x| >= 128
                const float rscale = *reinterpret_cast<volatile float *>(&cscale[it & 3]);
                if (warp_valid) {
                    // ldc is a compile-time constant: every store address is base + immediate
                    float *dst = C + (size_t)col0 * LDC + g * ROWS + m;
                    const bool full = (col0 + NT <= ncol);
#pragma unroll 1
                    for (int n0 = 0; n0 < NT; n0 += 32) {
                        float v[32];
                        tmem_ld32(lane_base + a * NT + n0, v);
                        tmem_ld_wait();
                        if (ok[g]) {
                            float *d0 = dst + (size_t)n0 * LDC;
                            if (full) {
#pragma unroll
                                for (int j = 0; j < 32; j++) d0[j * LDC] = fmaf(v[j], rscale, bg[g]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; j++)
                                    if (col0 + n0 + j < ncol) d0[j * LDC] = fmaf(v[j], rscale, bg[g]);
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acce[a]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, G::TCOLS);
}

template <int K, int ROWS, int NTILE, int NT, int LDC>
static int launch_affine_cfg(const float *X, int ncol, const uint8_t *wimg, const float *bias, int M, float *C,
                             const int *src_col, cudaStream_t s) {
    using G = GemmCfg<K, ROWS, NTILE, NT>;
    // persistent CTAs: one per SM unless SCRAPPIE_B200_AFFINE_CTAS caps it (read once) -- the kernel is HBM-bound and
    // holds a whole SM (216 KB of shared memory) for as long as it runs
    static const int max_ctas = [] { const char *e = getenv("SCRAPPIE_B200_AFFINE_CTAS"); const int v = e ? atoi(e) : 0; return (v > 0 && v < 148) ? v : 148; }();
    const int nchunk = (ncol + NT - 1) / NT;
    const int grid = nchunk < max_ctas ? nchunk : max_ctas;
    affine_tc_kernel<K, ROWS, NTILE, NT, LDC><<<grid, 416, G::SMEM, s>>>(X, ncol, wimg, bias, M, C, src_col);
    return 0;
}

// GRU input transform: M = 3H rows as three tiles (z, r, candidate) of H rows, K = H.
int launch_affine_tc(const float *X, int ncol, int H, const uint8_t *wimg, const float *bias, float *C, const int *src_col,
                     cudaStream_t s) {
    if (ncol <= 0) return 0;
    if (H == 96) return launch_affine_cfg<96, 96, 3, SB2_AFFINE_NT, 288>(X, ncol, wimg, bias, 3 * H, C, src_col, s);
    if (H == 112) return launch_affine_cfg<112, 112, 3, 64, 336>(X, ncol, wimg, bias, 3 * H, C, src_col, s);
    return -1;
}


// ---------------------------------------------------------------------------------
// output head: FF -> softmax (with temperature) -> robust log, one kernel
// ---------------------------------------------------------------------------------
// softmax_with_temperature + robustlog_activation_inplace (src/layers.c:340-357, :79-94) for the
// 4^k + 1 state transducer models with 1025 states:
//     p[m] = exp((b[m] + W[m] . (x / xdiv)) / cdiv) / sum_m' (...),   out = log(min_prob + (1 - min_prob) p)
// A chunk of NT = 64 columns keeps ALL 1024 k-mer logits in TMEM (8 tiles x 64 columns = the
// SM's 512 TMEM columns), so the column sum is local and the posterior is written to HBM exactly
// once, already normalised.  The 393 KB of split-fp16 weights do not fit in shared memory: a TMA
// bulk-copy ring streams the 8 tile images from L2 once per chunk.  The stay state (row 1024) is a
// 96-term dot product per column on the CUDA cores.
//
// warps 0-7  epilogue   (TMEM lane quarter q = warp % 4, column half ch = warp / 4)
//            pass A: e = exp(logit) written back to TMEM, column sums; pass B: normalise, log, store
// warp  8    UMMA issuer        warp 9  TMA weight streamer       warps 10-13  activation producers
template <int K>
struct HeadCfg {
// Weight ring depth.  Two stages (153 KB of shared memory in all) measure the same as three (202 KB) and leave room for a
// co-resident decode CTA (44 KB): the kernel then does not have to wait for a completely empty SM.
#ifndef SB2_HEAD_WSTAGES
#define SB2_HEAD_WSTAGES 2
#endif
    static constexpr int NT = 64, NTILE = 8, WSTAGES = SB2_HEAD_WSTAGES;
    static constexpr uint32_t LBO_A = 128, SBO_A = (K / 8) * 128;
    static constexpr uint32_t TILE_A = 128 * K * 2, WTILE = 2 * TILE_A;
    static constexpr uint32_t LBO_B = 16 * NT + 16, SBO_B = 128;
    static constexpr uint32_t TILE_B = (K / 8) * LBO_B, STAGE_B = 2 * TILE_B;
    static constexpr uint32_t PART = 2 * 2 * 4 * 32 * 4;                        // [buf][ch][q][lane] floats, x2 arrays
    static constexpr uint32_t OFF_B = WSTAGES * WTILE, OFF_PART = OFF_B + 2 * STAGE_B, OFF_BAR = OFF_PART + 2 * PART;
    static constexpr uint32_t SMEM = OFF_BAR + 32 * 8 + 16;
    static constexpr int NKS = K / 16;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

template <bool FAST>
__device__ __forceinline__ float head_exp(float x) {
    if (!FAST) return exp_cephes(x);
    x = fminf(fmaxf(x, -88.3762626647949f), 88.3762626647949f);         // the reference's clamp
    return ex2_approx(x * 1.4426950408889634f);
}
template <bool FAST>
__device__ __forceinline__ float head_log(float x) {
    if (!FAST) return log_cephes(x);
    return lg2_approx(x) * 0.6931471805599453f;
}

// EW = column slices of a chunk among the epilogue warps: 4 EW epilogue warps (TMEM lane quarter q = warp % 4,
// slice ch = warp / 4, CW = NT / EW columns each).  EW = 4 (16 warps, 16 columns each) keeps twice as many TMEM loads,
// stores and global loads in flight as EW = 2: the epilogue, not the tensor pipe or L2, sets this kernel's pace.
template <int CW>
__device__ __forceinline__ void tmem_ld_cw(uint32_t taddr, float (&v)[CW]) {
    if constexpr (CW == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
}
template <int CW>
__device__ __forceinline__ void tmem_st_cw(uint32_t taddr, const float (&v)[CW]) {
    if constexpr (CW == 32) tmem_st32(taddr, v); else tmem_st16(taddr, v);
}

template <int K, bool FAST, int EW>
__global__ void __launch_bounds__(32 * (4 * EW + 6), 1)
head_softmax_tc_kernel(const float *__restrict__ X, int ncol, const uint8_t *__restrict__ wimg,
                       const float *__restrict__ w_stay, const float *__restrict__ bias, float *__restrict__ post,
                       int ostride, float xdiv, float cdiv, float min_prob, int return_log) {
    using G = HeadCfg<K>;
    constexpr int NT = G::NT, NTILE = G::NTILE;
    constexpr int CW = NT / EW;                         // columns per epilogue warp
    constexpr int KSPLIT = 32 / CW;                     // lanes per column: they split the stay dot product
    constexpr int NEW = 4 * EW;                         // epilogue warps
    static_assert(CW == 32 || CW == 16, "column slice");
    // column stride on the device: 1056 floats = 33 lines of 128 bytes, so that a warp's store of 32 consecutive states
    // is ONE aligned line (with the reference's 1028 every column starts 16 bytes later than the one before and most
    // stores straddle two lines); the stay state and the three padding lanes sit at 1024..1027, 1028..1055 are unused
    constexpr int OSTRIDE = 1056;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *w_ring = smem;
    uint8_t *b_ring = smem + G::OFF_B;
    float *part_sum = reinterpret_cast<float *>(smem + G::OFF_PART);            // [2][EW][4][CW]
    float *part_stay = part_sum + 2 * EW * 4 * CW;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + G::OFF_BAR);
    uint64_t *full_b = bars, *empty_b = bars + 2, *wfull = bars + 4, *wempty = bars + 7, *tile_full = bars + 10,
             *tile_free = bars + 18;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 26);

    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int nchunk = (ncol + NT - 1) / NT;

    if (tid == 0) {
        for (int i = 0; i < 2; i++) { mbar_init(&full_b[i], 4); mbar_init(&empty_b[i], 1); }
        for (int i = 0; i < G::WSTAGES; i++) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
        for (int i = 0; i < NTILE; i++) { mbar_init(&tile_full[i], 1); mbar_init(&tile_free[i], NEW); }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == NEW) {
        // ---- UMMA issuer ------------------------------------------------------------------
        const uint32_t idesc = umma_idesc_f16(128, NT);
        constexpr uint64_t TA = G::TILE_A >> 4, KA = (2 * G::LBO_A) >> 4, KB = (2 * G::LBO_B) >> 4, TB = G::TILE_B >> 4;
        uint32_t it = 0, wslot = 0, wphase = 0;
        for (int c = blockIdx.x; c < nchunk; c += gridDim.x, it++) {
            const uint32_t s = it & 1;
            mbar_wait(&full_b[s], (it >> 1) & 1);
            const uint64_t dB = umma_desc(smem_u32(b_ring + s * G::STAGE_B), G::LBO_B, G::SBO_B);
#pragma unroll 1
            for (int t = 0; t < NTILE; t++) {
                mbar_wait(&wfull[wslot], wphase);
                mbar_wait(&tile_free[t], (it & 1) ^ 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t dcol = tmem + t * NT;
                    const uint64_t w_hi = umma_desc(smem_u32(w_ring + wslot * G::WTILE), G::LBO_A, G::SBO_A), w_lo = w_hi + TA;
#pragma unroll
                    for (int ks = 0; ks < G::NKS; ks++) umma_f16(dcol, w_lo + ks * KA, dB + ks * KB, idesc, ks > 0);
#pragma unroll
                    for (int ks = 0; ks < G::NKS; ks++) umma_f16(dcol, w_hi + ks * KA, dB + TB + ks * KB, idesc, 1);
#pragma unroll
                    for (int ks = 0; ks < G::NKS; ks++) umma_f16(dcol, w_hi + ks * KA, dB + ks * KB, idesc, 1);
                    umma_commit(&wempty[wslot]);
                    umma_commit(&tile_full[t]);
                }
                __syncwarp();
                if (++wslot == G::WSTAGES) { wslot = 0; wphase ^= 1; }
            }
            if (elect_one()) umma_commit(&empty_b[s]);
            __syncwarp();
        }
    } else if (warp == NEW + 1) {
        // ---- weight streamer: 8 tile images per chunk through a 3-deep TMA ring ------------
        uint32_t wslot = 0, wphase = 0;
        for (int c = blockIdx.x; c < nchunk; c += gridDim.x) {
#pragma unroll 1
            for (int t = 0; t < NTILE; t++) {
                mbar_wait(&wempty[wslot], wphase ^ 1);
                if (lane == 0) {
                    mbar_arrive_expect_tx(&wfull[wslot], G::WTILE);
                    uint8_t *dst = w_ring + wslot * G::WTILE;
                    const uint8_t *src = wimg + (size_t)t * G::WTILE;
                    bulk_g2s(dst, src, G::TILE_A, &wfull[wslot]);
                    bulk_g2s(dst + G::TILE_A, src + G::TILE_A, G::TILE_A, &wfull[wslot]);
                }
                __syncwarp();
                if (++wslot == G::WSTAGES) { wslot = 0; wphase ^= 1; }
            }
        }
    } else if (warp >= NEW + 2) {
        // ---- producers ------------------------------------------------------------------------
        const int pt = tid - 32 * (NEW + 2);
        constexpr int K8 = K / 8;
        constexpr int UNITS = NT * K8;
        const float scale = OPERAND_SCALE;
        uint32_t it = 0;
        for (int c = blockIdx.x; c < nchunk; c += gridDim.x, it++) {
            const uint32_t s = it & 1;
            mbar_wait(&empty_b[s], ((it >> 1) & 1) ^ 1);
            uint8_t *b_hi = b_ring + s * G::STAGE_B, *b_lo = b_hi + G::TILE_B;
            const int col0 = c * NT;
            const float *src = X + (size_t)col0 * K;
            const int nvalid = min(NT, ncol - col0) * K8;
#pragma unroll 2
            for (int u = pt; u < UNITS; u += 128) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                if (u < nvalid) {
                    a = *reinterpret_cast<const float4 *>(src + (size_t)u * 8);
                    b = *reinterpret_cast<const float4 *>(src + (size_t)u * 8 + 4);
                    if (xdiv != 1.0f) {
                        a.x /= xdiv; a.y /= xdiv; a.z /= xdiv; a.w /= xdiv;
                        b.x /= xdiv; b.y /= xdiv; b.z /= xdiv; b.w /= xdiv;
                    }
                }
                const int n = u / K8, k8 = u % K8;
                uint4 hi, lo;
                split8(a, b, scale, hi, lo);
                const uint32_t off = (uint32_t)(n >> 3) * G::SBO_B + (uint32_t)k8 * G::LBO_B + (uint32_t)(n & 7) * 16;
                *reinterpret_cast<uint4 *>(b_hi + off) = hi;
                *reinterpret_cast<uint4 *>(b_lo + off) = lo;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_b[s]);
        }
    } else {
        // ---- epilogue ---------------------------------------------------------------------------
        const int q = warp & 3, ch = warp >> 2;
        const int m = q * 32 + lane;
        const int cl = lane & (CW - 1);                   // this lane's column inside the slice (sums, stay row)
        const int kh = lane / CW;                         // which part of the stay dot product this lane takes
        const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + ch * CW;
        constexpr int KQ = K / 4;                         // the stay dot product is split over the 4 q-warps ...
        constexpr int KL = KQ / KSPLIT;                   // ... and over the KSPLIT lanes that share a column
        static_assert(KL % 4 == 0, "stay split");
        const float b_stay = bias[NTILE * 128];
        float ws[KL];
#pragma unroll
        for (int i = 0; i < KL; i++) ws[i] = w_stay[q * KQ + kh * KL + i];
        const float keep = 1.0f - min_prob;
        // FAST: exp((acc * 2^-16 + b) / cdiv) = 2^(acc * sc + b * bsc), one FFMA in front of ex2.approx
        const float inv_cdiv = 1.0f / cdiv;
        const float sc = RESULT_SCALE * inv_cdiv * 1.4426950408889634f, bsc = inv_cdiv * 1.4426950408889634f;
        float4 xs4[KL / 4];
        auto load_stay_x = [&](int chunk) {
            const int col = min(chunk * NT + ch * CW + cl, ncol - 1);      // clamped: out-of-range columns are never stored
            const float4 *xp = reinterpret_cast<const float4 *>(X + (size_t)col * K + q * KQ + kh * KL);
#pragma unroll
            for (int i = 0; i < KL / 4; i++) xs4[i] = __ldg(xp + i);
        };
        load_stay_x(blockIdx.x);
        uint32_t it = 0;
        for (int c = blockIdx.x; c < nchunk; c += gridDim.x, it++) {
            const int col0 = c * NT + ch * CW;            // first column of this warp's slice
            const int mycol = col0 + cl;
            // stay logit, partial over this lane's part of k for column `mycol`: the activations were fetched while
            // the previous chunk was being finished (xs4), the next chunk's are requested right away
            float sp = 0.0f;
#pragma unroll
            for (int i = 0; i < KL / 4; i++) {
                float4 x4 = xs4[i];
                if (xdiv != 1.0f) { x4.x /= xdiv; x4.y /= xdiv; x4.z /= xdiv; x4.w /= xdiv; }
                sp = fmaf(ws[4 * i], x4.x, sp); sp = fmaf(ws[4 * i + 1], x4.y, sp);
                sp = fmaf(ws[4 * i + 2], x4.z, sp); sp = fmaf(ws[4 * i + 3], x4.w, sp);
            }
            if (KSPLIT == 2) sp += __shfl_xor_sync(0xffffffffu, sp, 16);
            load_stay_x(c + (int)gridDim.x);
            // pass A: e = exp(logit), column sums over this thread's 8 rows
            float cs[CW];
#pragma unroll
            for (int j = 0; j < CW; j++) cs[j] = 0.0f;
#pragma unroll 1
            for (int t = 0; t < NTILE; t++) {
                mbar_wait(&tile_full[t], it & 1);
                tc_fence_after();
                float v[CW];
                tmem_ld_cw<CW>(tbase + t * NT, v);
                tmem_ld_wait();
                const float bt = __ldg(bias + t * 128 + m);
                if (FAST) {
                    const float bt2 = bt * bsc;
#pragma unroll
                    for (int j = 0; j < CW; j++) {
                        // clamp = the reference's exp_ps input clamp (+-88.376) expressed in base 2
                        const float e = ex2_approx(fminf(fmaxf(fmaf(v[j], sc, bt2), -127.5f), 127.5f));
                        cs[j] += e;
                        v[j] = e;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < CW; j++) {
                        const float e = exp_cephes((fmaf(v[j], RESULT_SCALE, bt)) / cdiv);
                        cs[j] += e;
                        v[j] = e;
                    }
                }
                tmem_st_cw<CW>(tbase + t * NT, v);
            }
            tmem_st_wait();
            // column sums over the 32 lanes (= 32 rows): halving butterfly over the CW columns, then (CW = 16) the two
            // half-warps are added; lane L ends with the sum of column L % CW
#pragma unroll
            for (int w = CW / 2; w > 0; w >>= 1) {
                const bool up = (lane & w) != 0;
#pragma unroll
                for (int j = 0; j < w; j++) {
                    const float lo_v = cs[j], hi_v = cs[j + w];
                    const float send = up ? lo_v : hi_v;
                    const float mine_v = up ? hi_v : lo_v;
                    cs[j] = mine_v + __shfl_xor_sync(0xffffffffu, send, w);
                }
            }
            if (KSPLIT == 2) cs[0] += __shfl_xor_sync(0xffffffffu, cs[0], 16);
            const int pb = ((it & 1) * EW + ch) * 4 * CW;
            if (lane < CW) {
                part_sum[pb + q * CW + cl] = cs[0];
                part_stay[pb + q * CW + cl] = sp;
            }
            named_bar_sync(1 + ch, 128);
            float tot = 0.0f, sl = 0.0f;
#pragma unroll
            for (int qq = 0; qq < 4; qq++) { tot += part_sum[pb + qq * CW + cl]; sl += part_stay[pb + qq * CW + cl]; }
            const float e_stay = FAST ? ex2_approx(fminf(fmaxf((b_stay + sl) * bsc, -127.5f), 127.5f))
                                      : exp_cephes((b_stay + sl) / cdiv);
            const float recip = __fdiv_rn(1.0f, tot + e_stay);
            if (q == 0 && lane < CW && mycol < ncol) {
                // stay state and the three padding lanes (exp(0) = 1 in the reference, normalised like the rest)
                float ps = e_stay * recip, pp = recip;
                if (return_log) { ps = head_log<FAST>(min_prob + keep * ps); pp = head_log<FAST>(min_prob + keep * pp); }
                float *o = post + (size_t)mycol * OSTRIDE + NTILE * 128;
                o[0] = ps; o[1] = pp; o[2] = pp; o[3] = pp;
            }
            // pass B: normalise, robust log, store.  out = log(min_prob + keep * e * recip)
            const float krc_mine = return_log ? keep * recip : recip;
            float rc[CW];
#pragma unroll
            for (int j = 0; j < CW; j++) rc[j] = __shfl_sync(0xffffffffu, krc_mine, j);
            const int nok = min(CW, ncol - col0);
#pragma unroll 1
            for (int t = 0; t < NTILE; t++) {
                float v[CW];
                tmem_ld_cw<CW>(tbase + t * NT, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tile_free[t]);
                float *dst = post + (size_t)col0 * OSTRIDE + t * 128 + m;
                if (return_log) {
#pragma unroll
                    for (int j = 0; j < CW; j++) v[j] = head_log<FAST>(fmaf(v[j], rc[j], min_prob));
                } else {
#pragma unroll
                    for (int j = 0; j < CW; j++) v[j] = v[j] * rc[j];
                }
                if (nok == CW) {
#pragma unroll
                    for (int j = 0; j < CW; j++) dst[j * OSTRIDE] = v[j];
                } else {
#pragma unroll
                    for (int j = 0; j < CW; j++)
                        if (j < nok) dst[j * OSTRIDE] = v[j];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

size_t head_image_bytes(int K) { return gemm_image_bytes(8, 128, K); }

int launch_head_softmax_tc(const float *X, int ncol, int K, const uint8_t *wimg, const float *w_stay, const float *bias,
                           float *post, int ostride, float xdiv, float cdiv, float min_prob, int return_log,
                           int exact_math, cudaStream_t s) {
    if (ncol <= 0) return 0;
    if (K != 96 || ostride != 1056) return -1;
    using G = HeadCfg<96>;
    static const int ew = [] { const char *e = getenv("SCRAPPIE_B200_HEAD_SLICES"); return (e && atoi(e) == 2) ? 2 : 4; }();
    static const int max_ctas = [] { const char *e = getenv("SCRAPPIE_B200_HEAD_CTAS"); const int v = e ? atoi(e) : 0; return (v > 0 && v < 148) ? v : 148; }();
    const int nchunk = (ncol + G::NT - 1) / G::NT;
    const int grid = nchunk < max_ctas ? nchunk : max_ctas;
    if (exact_math)
        head_softmax_tc_kernel<96, false, 2><<<grid, 448, G::SMEM, s>>>(X, ncol, wimg, w_stay, bias, post, ostride, xdiv, cdiv, min_prob, return_log);
    else if (ew == 2)
        head_softmax_tc_kernel<96, true, 2><<<grid, 448, G::SMEM, s>>>(X, ncol, wimg, w_stay, bias, post, ostride, xdiv, cdiv, min_prob, return_log);
    else
        head_softmax_tc_kernel<96, true, 4><<<grid, 704, G::SMEM, s>>>(X, ncol, wimg, w_stay, bias, post, ostride, xdiv, cdiv, min_prob, return_log);
    return 0;
}

// Per-device function attributes of this file's kernels (called once per engine, after cudaSetDevice): the
// attribute belongs to the device, so a process-wide "configured" flag would leave a second GPU unconfigured.
int configure_gemm_kernels() {
    const cudaFuncAttribute A = cudaFuncAttributeMaxDynamicSharedMemorySize;
    bool ok = cudaFuncSetAttribute(affine_tc_kernel<96, 96, 3, SB2_AFFINE_NT, 288>, A, (int)GemmCfg<96, 96, 3, SB2_AFFINE_NT>::SMEM) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(affine_tc_kernel<112, 112, 3, 64, 336>, A, (int)GemmCfg<112, 112, 3, 64>::SMEM) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(head_softmax_tc_kernel<96, true, 2>, A, (int)HeadCfg<96>::SMEM) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(head_softmax_tc_kernel<96, true, 4>, A, (int)HeadCfg<96>::SMEM) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(head_softmax_tc_kernel<96, false, 2>, A, (int)HeadCfg<96>::SMEM) == cudaSuccess;
    return ok ? 0 : -1;
}

}  // namespace sb2
