// Device transcendental functions.
//
// sb2_exp / sb2_log evaluate the same cephes polynomials, in the same operation order
// and without FMA contraction, as the reference's exp_ps / log_ps
// (src/sse_mathfun.h:123-290), so that given equal inputs the results are bit-identical
// to the CPU path (input clamp +-88.376 included).  logistic / tanh / elu follow
// src/util.h:180-198: sigma(x) = 1 / (1 + exp(-x)), tanh(x) = 2 sigma(2x) - 1,
// elu(x) = x >= 0 ? x : exp(x) - 1.
//
// The *_fast variants use the SFU (ex2.approx / rcp.approx) and are used only inside
// the latency-critical GRU scan; their error (a few ulp) is measured against the oracle
// in tests/test_gpu_parity.py.
#pragma once
#include <cuda_runtime.h>

namespace sb2 {

__device__ __forceinline__ float exp_cephes(float x) {
    x = fminf(x, 88.3762626647949f);
    x = fmaxf(x, -88.3762626647949f);
    float fx = __fadd_rn(__fmul_rn(x, 1.44269504088896341f), 0.5f);
    float fl = (float)(int)fx;                  // truncation, then corrected to floor
    if (fl > fx) fl = __fsub_rn(fl, 1.0f);
    fx = fl;
    const float t = __fmul_rn(fx, 0.693359375f);
    float z = __fmul_rn(fx, -2.12194440e-4f);
    x = __fsub_rn(x, t);
    x = __fsub_rn(x, z);
    z = __fmul_rn(x, x);
    float y = 1.9875691500E-4f;
    y = __fadd_rn(__fmul_rn(y, x), 1.3981999507E-3f);
    y = __fadd_rn(__fmul_rn(y, x), 8.3334519073E-3f);
    y = __fadd_rn(__fmul_rn(y, x), 4.1665795894E-2f);
    y = __fadd_rn(__fmul_rn(y, x), 1.6666665459E-1f);
    y = __fadd_rn(__fmul_rn(y, x), 5.0000001201E-1f);
    y = __fmul_rn(y, z);
    y = __fadd_rn(y, x);
    y = __fadd_rn(y, 1.0f);
    const float pow2n = __int_as_float(((int)fx + 127) << 23);
    return __fmul_rn(y, pow2n);
}

__device__ __forceinline__ float log_cephes(float x) {
    const bool invalid = (x <= 0.0f);
    x = fmaxf(x, __int_as_float(0x00800000));
    int e_i = (int)(__float_as_uint(x) >> 23) - 127;
    x = __uint_as_float((__float_as_uint(x) & 0x807fffffu) | 0x3f000000u);
    float e = __fadd_rn((float)e_i, 1.0f);
    const bool small = (x < 0.707106781186547524f);
    const float keep = small ? x : 0.0f;
    x = __fsub_rn(x, 1.0f);
    e = __fsub_rn(e, small ? 1.0f : 0.0f);
    x = __fadd_rn(x, keep);
    const float z = __fmul_rn(x, x);
    float y = 7.0376836292E-2f;
    y = __fadd_rn(__fmul_rn(y, x), -1.1514610310E-1f);
    y = __fadd_rn(__fmul_rn(y, x), 1.1676998740E-1f);
    y = __fadd_rn(__fmul_rn(y, x), -1.2420140846E-1f);
    y = __fadd_rn(__fmul_rn(y, x), 1.4249322787E-1f);
    y = __fadd_rn(__fmul_rn(y, x), -1.6668057665E-1f);
    y = __fadd_rn(__fmul_rn(y, x), 2.0000714765E-1f);
    y = __fadd_rn(__fmul_rn(y, x), -2.4999993993E-1f);
    y = __fadd_rn(__fmul_rn(y, x), 3.3333331174E-1f);
    y = __fmul_rn(y, x);
    y = __fmul_rn(y, z);
    y = __fadd_rn(y, __fmul_rn(e, -2.12194440e-4f));
    y = __fsub_rn(y, __fmul_rn(z, 0.5f));
    x = __fadd_rn(x, y);
    x = __fadd_rn(x, __fmul_rn(e, 0.693359375f));
    return invalid ? __int_as_float(0x7fc00000) : x;
}

__device__ __forceinline__ float logistic_cephes(float x) {
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, exp_cephes(-x)));
}
__device__ __forceinline__ float tanh_cephes(float x) {
    const float y = logistic_cephes(__fadd_rn(x, x));
    return __fsub_rn(__fadd_rn(y, y), 1.0f);
}
__device__ __forceinline__ float elu_cephes(float x) {
    return (x >= 0.0f) ? x : __fsub_rn(exp_cephes(x), 1.0f);
}

// SFU versions for the scan's critical path.
__device__ __forceinline__ float logistic_fast(float x) {
    // 1 / (1 + 2^(-x * log2 e)); ex2.approx saturates cleanly to 0 / +inf
    const float e = exp2f(-1.4426950408889634f * x);
    return __fdividef(1.0f, 1.0f + e);
}
__device__ __forceinline__ float tanh_fast(float x) {
    const float y = logistic_fast(x + x);
    return (y + y) - 1.0f;
}

// libm-style log-sum-exp used by the CRF partition function (src/util.h:162-164)
__device__ __forceinline__ float logsumexp2(float a, float b) {
    return fmaxf(a, b) + log1pf(expf(-fabsf(a - b)));
}

}  // namespace sb2
