// Micro-benchmark: latency / throughput of small tcgen05.mma (M=128, N=16, K=16, fp16, A in TMEM)
// and the arrival time of tcgen05.commit when more MMAs follow it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../scrappie_b200/csrc -o mma_probe mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace sb2::tc;

// mode bit0: 0 = every MMA accumulates into the same D columns, 1 = MMA i uses D columns 16*(i%8)
// mode bit1: 0 = A from TMEM, 1 = A from shared memory
template <int n1, int n2, int mode, int N>
__global__ void __launch_bounds__(288, 1) probe(int reps, int nwait, long long *out) {
    __shared__ __align__(128) uint8_t bop[16 * 1024];
    __shared__ __align__(128) uint8_t aop[30 * 1024];
    __shared__ __align__(8) uint64_t bars[3];
    __shared__ uint32_t slot;
    __shared__ long long t0s;
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    for (int i = tid; i < 16 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(bop)[i] = 0x3c003c00u;
    for (int i = tid; i < 30 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(aop)[i] = 0x3c003c00u;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t LBO_B = 16u * N + 16u;
    const uint32_t idesc = umma_idesc_f16(128, N);
    const uint64_t dB = umma_desc(smem_u32(bop), LBO_B, 128);
    const uint64_t dA = umma_desc(smem_u32(aop), 128, 12 * 128);
    const uint64_t KB = (2 * LBO_B) >> 4;
    long long sumA = 0, sumB = 0, sumI = 0;
    for (int rep = 0; rep < reps; rep++) {
        __syncthreads();
        if (warp == 8) {
            const long long t0 = clock64();
            if (lane == 0) t0s = t0;
            if (elect_one()) {
#pragma unroll
                for (int i = 0; i < n1; i++) {
                    const uint32_t dcol = tmem + 384 + ((mode & 1) ? 16 * (i % 8) : 0);
                    if (mode & 2) umma_f16(dcol, dA + (i % 6) * 16, dB + (i % 6) * KB, idesc, 1);
                    else umma_f16_ts(dcol, tmem + (i % 12) * 8, dB + (i % 6) * KB, idesc, 1);
                }
                umma_commit(&bars[0]);
#pragma unroll
                for (int i = 0; i < n2; i++) {
                    const uint32_t dcol = tmem + 384 + ((mode & 1) ? 16 * (i % 8) : 0);
                    if (mode & 2) umma_f16(dcol, dA + (i % 6) * 16, dB + (i % 6) * KB, idesc, 1);
                    else umma_f16_ts(dcol, tmem + 96 + (i % 12) * 8, dB + (i % 6) * KB, idesc, 1);
                }
                umma_commit(&bars[1]);
            }
            __syncwarp();
            sumI += clock64() - t0;
        } else if (warp >= nwait) {
            // idle
        } else if ((warp & 1) == 0) {
            mbar_wait(&bars[0], rep & 1);
            const long long t = clock64();
            __syncwarp();
            // t0s was written before the MMAs were issued, long before this barrier completes
            sumA += t - *(volatile long long *)&t0s;
        } else {
            mbar_wait(&bars[1], rep & 1);
            const long long t = clock64();
            __syncwarp();
            sumB += t - *(volatile long long *)&t0s;
        }
    }
    __syncthreads();
    if (tid == 0) out[0] = sumA;
    if (tid == 32) out[1] = sumB;
    if (tid == 256) out[2] = sumI;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int n1, int n2, int mode, int N>
void run(long long *d, int reps, int nwait = 2) {
    cudaMemset(d, 0, 64);
    probe<n1, n2, mode, N><<<1, 288>>>(reps, nwait, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    long long h[3];
    cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("%4d %4d %4d %4d w%d | %8.1f %8.1f %8.1f\n", n1, n2, mode, N, nwait, (double)h[0] / reps, (double)h[1] / reps, (double)h[2] / reps);
}

int main() {
    long long *d;
    cudaMalloc(&d, 64);
    const int reps = 200;
    printf("%4s %4s %4s %4s | %8s %8s %8s   (cycles from issue start; A = first commit, B = second commit, I = issue loop)\n",
           "n1", "n2", "mode", "N", "A", "B", "I");
    run<1, 0, 0, 16>(d, reps); run<12, 0, 0, 16>(d, reps); run<24, 0, 0, 16>(d, reps); run<12, 12, 0, 16>(d, reps);
    run<12, 0, 0, 16>(d, reps, 4); run<12, 0, 0, 16>(d, reps, 8); run<12, 12, 0, 16>(d, reps, 8); run<24, 0, 0, 16>(d, reps, 8);
    run<12, 0, 0, 32>(d, reps); run<12, 0, 0, 64>(d, reps);
    run<12, 0, 2, 16>(d, reps); run<12, 12, 2, 16>(d, reps); run<12, 0, 2, 64>(d, reps);
    return 0;
}
