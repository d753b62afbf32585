// Events (LSTM) model kernels: window() and lstm_forward / lstm_backward (src/layers.c:119-146, :673-832).
//
// This is the legacy model behind interface/scrappie.h's nanonet_posterior; it runs in fp32 on the CUDA cores
// with the cephes-identical gate functions of device_math.cuh (the GRU models are the tensor-core path).
//   lstm_scan_kernel: one CTA = 8 reads stepping together, 4H threads.  Thread k keeps row k of sW (the weights
//   into gate unit k) in registers and computes  pre[k] = x[k] + sW[k] . h  for the 8 reads from the state held
//   in shared memory as [H][8]; threads j < H then evaluate the gates of hidden unit j (cell state in registers):
//       forget = sigma(pre[2H+j] + c p[H+j]) c        update = sigma(pre[H+j] + c p[j]) tanh(pre[j])
//       c'     = forget + update                      out    = sigma(pre[3H+j] + c' p[2H+j]) tanh(c')
#include "device_math.cuh"
#include "kernels.h"

namespace sb2 {
namespace {

constexpr int LSTM_R = 8;

template <int H>
__global__ void __launch_bounds__(4 * H, 1)
lstm_scan_kernel(const float *__restrict__ Xin, const float *__restrict__ sW, const float *__restrict__ peep,
                 float *__restrict__ out, BatchDims d, int backward) {
    __shared__ __align__(16) float hs[2][H][LSTM_R];
    __shared__ __align__(16) float pre[4 * H][LSTM_R];
    __shared__ int s_T[LSTM_R], s_col[LSTM_R];

    const int k = threadIdx.x;
    const int r0 = blockIdx.x * LSTM_R;
    if (k < LSTM_R) {
        const int r = r0 + k;
        s_T[k] = (r < d.nread) ? d.nblock[r] : 0;
        s_col[k] = (r < d.nread) ? d.col_off[r] : 0;
    }
    for (int i = k; i < 2 * H * LSTM_R; i += 4 * H) (&hs[0][0][0])[i] = 0.0f;

    float w[H];
    {
        const float *row = sW + (size_t)k * H;
#pragma unroll
        for (int i = 0; i < H; i += 4) {
            const float4 v = *reinterpret_cast<const float4 *>(row + i);
            w[i] = v.x; w[i + 1] = v.y; w[i + 2] = v.z; w[i + 3] = v.w;
        }
    }
    const int j = (k < H) ? k : 0;
    const float p_in = peep[j], p_forget = peep[H + j], p_out = peep[2 * H + j];
    __syncthreads();
    int T[LSTM_R], col[LSTM_R], Tmax = 0;
#pragma unroll
    for (int r = 0; r < LSTM_R; r++) { T[r] = s_T[r]; col[r] = s_col[r]; Tmax = max(Tmax, T[r]); }

    float cell[LSTM_R], xn[LSTM_R];
#pragma unroll
    for (int r = 0; r < LSTM_R; r++) {
        cell[r] = 0.0f;
        const int t = backward ? (T[r] - 1) : 0;
        xn[r] = (T[r] > 0) ? Xin[(size_t)(col[r] + t) * (4 * H) + k] : 0.0f;
    }

    for (int s = 0; s < Tmax; s++) {
        const int cur = s & 1;
        float acc[LSTM_R];
#pragma unroll
        for (int r = 0; r < LSTM_R; r++) acc[r] = 0.0f;
#pragma unroll
        for (int i = 0; i < H; i++) {
            const float4 ha = *reinterpret_cast<const float4 *>(&hs[cur][i][0]);
            const float4 hb = *reinterpret_cast<const float4 *>(&hs[cur][i][4]);
            acc[0] = fmaf(w[i], ha.x, acc[0]); acc[1] = fmaf(w[i], ha.y, acc[1]);
            acc[2] = fmaf(w[i], ha.z, acc[2]); acc[3] = fmaf(w[i], ha.w, acc[3]);
            acc[4] = fmaf(w[i], hb.x, acc[4]); acc[5] = fmaf(w[i], hb.y, acc[5]);
            acc[6] = fmaf(w[i], hb.z, acc[6]); acc[7] = fmaf(w[i], hb.w, acc[7]);
        }
        *reinterpret_cast<float4 *>(&pre[k][0]) = make_float4(xn[0] + acc[0], xn[1] + acc[1], xn[2] + acc[2], xn[3] + acc[3]);
        *reinterpret_cast<float4 *>(&pre[k][4]) = make_float4(xn[4] + acc[4], xn[5] + acc[5], xn[6] + acc[6], xn[7] + acc[7]);
        if (s + 1 < Tmax) {
#pragma unroll
            for (int r = 0; r < LSTM_R; r++) {
                const int t = backward ? (T[r] - 2 - s) : (s + 1);
                xn[r] = (s + 1 < T[r]) ? Xin[(size_t)(col[r] + t) * (4 * H) + k] : 0.0f;
            }
        }
        __syncthreads();
        if (k < H) {
#pragma unroll
            for (int r = 0; r < LSTM_R; r++) {
                const float c = cell[r];
                const float forget = __fmul_rn(logistic_cephes(__fadd_rn(pre[2 * H + k][r], __fmul_rn(c, p_forget))), c);
                const float update = __fmul_rn(logistic_cephes(__fadd_rn(pre[H + k][r], __fmul_rn(c, p_in))), tanh_cephes(pre[k][r]));
                const float cn = __fadd_rn(forget, update);
                const float o = __fmul_rn(logistic_cephes(__fadd_rn(pre[3 * H + k][r], __fmul_rn(cn, p_out))), tanh_cephes(cn));
                cell[r] = cn;
                hs[cur ^ 1][k][r] = o;
                if (s < T[r]) {
                    const int t = backward ? (T[r] - 1 - s) : s;
                    out[(size_t)(col[r] + t) * H + k] = o;
                }
            }
        }
        __syncthreads();
    }
}

// window(features, 3, 1), src/layers.c:119-146, one CTA per read.  Column c of a read is
// [f[c-1], f[c], f[c+1]] with zeros past the end -- and column 0 is ALL ZERO: the reference's loop bound
// compares a negative int (w1 = -1) with a size_t, i.e. as a huge unsigned value, and never runs for column 0.
__global__ void window3_kernel(const float *__restrict__ feat, float *__restrict__ out, BatchDims d) {
    const int r = blockIdx.x;
    const int T = d.nblock[r];
    const float4 *f = reinterpret_cast<const float4 *>(feat) + d.col_off[r];
    float4 *o = reinterpret_cast<float4 *>(out) + (size_t)d.col_off[r] * 3;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = threadIdx.x; i < 3 * T; i += blockDim.x) {
        const int c = i / 3, w = i % 3 - 1;
        const int src = c + w;
        o[i] = (c == 0 || src >= T) ? zero : f[src];
    }
}

}  // namespace

void launch_window3(const float *feat, float *out, const BatchDims &d, cudaStream_t s) {
    if (d.nread > 0) window3_kernel<<<d.nread, 256, 0, s>>>(feat, out, d);
}

int launch_lstm_scan(const float *Xin, const float *sW, const float *peep, float *out, const BatchDims &d, int H,
                     int backward, cudaStream_t s) {
    if (H != 96) return -1;
    const int grid = (d.nread + LSTM_R - 1) / LSTM_R;
    lstm_scan_kernel<96><<<grid, 4 * 96, 0, s>>>(Xin, sW, peep, out, d, backward);
    return 0;
}

}  // namespace sb2
