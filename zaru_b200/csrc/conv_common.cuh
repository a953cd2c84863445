// Device helpers shared by kernels_conv.cu and kernels_thin.cu: activations, vector loads, residual reads.
#pragma once
#include <cuda_runtime.h>

#include "kernels.h"

namespace zb {
namespace {

__device__ __forceinline__ float apply_act(float v, const ActDev &a, int n) {
    switch (a.kind) {
        case ACT_RELU: return fmaxf(v, 0.0f);
        case ACT_PRELU: return v < 0.0f ? v * __ldg(a.slope + n) : v;
        case ACT_CLIP: return fminf(fmaxf(v, a.lo), a.hi);
        case ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
        default: return v;
    }
}

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// 16-byte asynchronous global->shared copy (LDGSTS); src_bytes = 0 zero-fills the destination.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, int src_bytes) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Packed FP32 FMA (Blackwell FFMA2: two fused multiply-adds per issued instruction; each lane rounds exactly
// like fmaf, so results are bit-identical to the scalar form while the FMA issue slots are halved).
// Measured (round 1): neutral in the thin kernel, SLOWER in the stem and GEMM-tile kernels (the (a,a) broadcast
// pairs cost extra moves and registers), so only the thin kernel uses it.
__device__ __forceinline__ void fma4(float4 &acc, const float4 &x, const float4 &w) {          // acc += x * w
    const float2 lo = __ffma2_rn(make_float2(x.x, x.y), make_float2(w.x, w.y), make_float2(acc.x, acc.y));
    const float2 hi = __ffma2_rn(make_float2(x.z, x.w), make_float2(w.z, w.w), make_float2(acc.z, acc.w));
    acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ void fma4s(float2 &acc_lo, float2 &acc_hi, float a, const float4 &w) {   // acc += a * w
    const float2 aa = make_float2(a, a);
    acc_lo = __ffma2_rn(aa, make_float2(w.x, w.y), acc_lo);
    acc_hi = __ffma2_rn(aa, make_float2(w.z, w.w), acc_hi);
}

// Activation on 4 consecutive channels starting at n (n % 4 == 0): ONE uniform branch per group of four
// instead of a switch per element (the per-element switch tripled the instruction count of the thin layers).
__device__ __forceinline__ void act4(float (&v)[4], const ActDev &a, int n) {
    if (a.kind == ACT_NONE) return;
    if (a.kind == ACT_RELU) {
#pragma unroll
        for (int q = 0; q < 4; q++) v[q] = fmaxf(v[q], 0.0f);
    } else if (a.kind == ACT_PRELU) {
        const float4 s = ldg4(a.slope + n);
        v[0] = v[0] < 0.0f ? v[0] * s.x : v[0];
        v[1] = v[1] < 0.0f ? v[1] * s.y : v[1];
        v[2] = v[2] < 0.0f ? v[2] * s.z : v[2];
        v[3] = v[3] < 0.0f ? v[3] * s.w : v[3];
    } else if (a.kind == ACT_CLIP) {
#pragma unroll
        for (int q = 0; q < 4; q++) v[q] = fminf(fmaxf(v[q], a.lo), a.hi);
    } else {
#pragma unroll
        for (int q = 0; q < 4; q++) v[q] = 1.0f / (1.0f + expf(-v[q]));
    }
}
__device__ __forceinline__ void act4(float4 &v, const ActDev &a, int n) {
    if (a.kind == ACT_NONE) return;
    float t[4] = {v.x, v.y, v.z, v.w};
    act4(t, a, n);
    v = make_float4(t[0], t[1], t[2], t[3]);
}

// Residual value for output pixel (img, oy, ox), channel n (n % 4 == 0 when vectorised by the caller).
__device__ __forceinline__ float residual_at(const EpiDev &e, int img, int oy, int ox, int n) {
    if (n >= e.res_Cs) return 0.0f;
    const float *base = e.res + (long long)img * e.res_img_stride;
    if (!e.res_pool) return __ldg(base + ((long long)oy * e.res_W + ox) * e.res_Cs + n);
    const float *p = base + ((long long)(2 * oy) * e.res_W + 2 * ox) * e.res_Cs + n;
    float a = __ldg(p), b = __ldg(p + e.res_Cs);
    float c = __ldg(p + (long long)e.res_W * e.res_Cs), d = __ldg(p + (long long)e.res_W * e.res_Cs + e.res_Cs);
    return fmaxf(fmaxf(a, b), fmaxf(c, d));
}

__device__ __forceinline__ float4 residual4_at(const EpiDev &e, int img, int oy, int ox, int n) {
    if (n >= e.res_Cs) return make_float4(0.f, 0.f, 0.f, 0.f);
    const float *base = e.res + (long long)img * e.res_img_stride;
    if (!e.res_pool) return ldg4(base + ((long long)oy * e.res_W + ox) * e.res_Cs + n);
    const float *p = base + ((long long)(2 * oy) * e.res_W + 2 * ox) * e.res_Cs + n;
    float4 a = ldg4(p), b = ldg4(p + e.res_Cs);
    float4 c = ldg4(p + (long long)e.res_W * e.res_Cs), d = ldg4(p + (long long)e.res_W * e.res_Cs + e.res_Cs);
    return make_float4(fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y)),
                       fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z)), fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w)));
}


// Depthwise KSxKS over channels [k, k+4) for one output pixel.  All loads of a row (KS=5) or of the whole
// window (KS=3) are issued before the FMAs that consume them, so one thread keeps 9 (5) independent
// 128-bit loads in flight instead of a dependent load->FMA chain.
template <int KS>
__device__ __forceinline__ float4 dw_window(const ConvDev &p, const float *__restrict__ a_base, int iy0, int ix0, int k) {
    float4 v = ldg4(p.dw_b + k);
    if (KS == 3) {
        float4 x[9];
#pragma unroll
        for (int ky = 0; ky < 3; ky++) {
            const int iy = iy0 + ky;
            const bool rowok = iy >= 0 && iy < p.H;
#pragma unroll
            for (int kx = 0; kx < 3; kx++) {
                const int ix = ix0 + kx;
                x[ky * 3 + kx] = (rowok && ix >= 0 && ix < p.W) ? ldg4(a_base + ((long long)iy * p.W + ix) * p.Cs_in + k)
                                                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int t = 0; t < 9; t++) {
            const float4 wv = ldg4(p.dw_w + t * p.Cs_in + k);
            v.x = fmaf(x[t].x, wv.x, v.x);
            v.y = fmaf(x[t].y, wv.y, v.y);
            v.z = fmaf(x[t].z, wv.z, v.z);
            v.w = fmaf(x[t].w, wv.w, v.w);
        }
    } else {
#pragma unroll 1
        for (int ky = 0; ky < KS; ky++) {
            const int iy = iy0 + ky;
            if (iy < 0 || iy >= p.H) continue;
            float4 x[KS];
#pragma unroll
            for (int kx = 0; kx < KS; kx++) {
                const int ix = ix0 + kx;
                x[kx] = (ix >= 0 && ix < p.W) ? ldg4(a_base + ((long long)iy * p.W + ix) * p.Cs_in + k)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int kx = 0; kx < KS; kx++) {
                const float4 wv = ldg4(p.dw_w + (ky * KS + kx) * p.Cs_in + k);
                v.x = fmaf(x[kx].x, wv.x, v.x);
                v.y = fmaf(x[kx].y, wv.y, v.y);
                v.z = fmaf(x[kx].z, wv.z, v.z);
                v.w = fmaf(x[kx].w, wv.w, v.w);
            }
        }
    }
    return v;
}


}  // namespace
}  // namespace zb
