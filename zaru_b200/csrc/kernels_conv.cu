// sm_100a kernels for the CNN forward pass (NHWC f32, FMA allowed).
//
//  conv_gemm_kernel  — every full convolution (stems 5x5/3x3 s2, 2x2 s2, 1x1 pointwise, dense heads,
//                      Gemm) as an implicit GEMM: M = images*Ho*Wo pixels, N = Cout, K = kh*kw*Cs_in.
//                      A tile gathered (or, for OP_DWPW, COMPUTED by the depthwise stage on the fly so the
//                      depthwise output never touches HBM), W tile staged in shared memory, register-tiled
//                      FFMA micro-kernel, fused epilogue: bias -> act1 -> (+ residual, optionally read
//                      through a 2x2 max-pool and zero channel-pad) -> act2 -> NHWC / head-layout store.
//  dw_kernel         — standalone depthwise kxk.
//  small ops         — 2x2 max-pool, bilinear x2 resize (half_pixel), global average pool, eltwise add/act,
//                      NCHW -> NHWC4 input conversion.
//
// Replaces the engine call `ort.session.run` / `plan.run` (crates/zaru/src/nn/mod.rs:496, :528).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <stdexcept>

#include "conv_common.cuh"
#include "kernels.h"

namespace zb {

std::atomic<long long> g_launch_count{0};
thread_local bool t_pdl_suppress = false;
bool pdl_enabled(int family) {
    static const int mask = getenv("ZB_PDL") ? atoi(getenv("ZB_PDL")) : 63;
    return (mask & family) != 0;
}
thread_local const char *t_kernel_name = nullptr;

bool launch_dwpw_thin(const ConvDev &p, cudaStream_t s);   // kernels_thin.cu

namespace {

// ------------------------------------------------------------------------------------------------
// Implicit-GEMM convolution.  MODE: 0 gather (any kh,kw,stride,pads), 1 pointwise (1x1/s1/p0),
// 2 fused depthwise producer.
// ------------------------------------------------------------------------------------------------

template <int BM, int BN, int TM, int TN, int BK, int MODE, int KS>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) conv_gemm_kernel(const ConvDev p) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int TXN = BN / TN;
    constexpr int QP = BK / 4;                          // k-quads (float4) per pixel per chunk
    constexpr int A_PER_THREAD = (BM * QP) / NT;        // float4 A elements per thread per chunk
    constexpr int W_TOTAL = BK * BN / 4;
    constexpr int W_PER_THREAD = (W_TOTAL + NT - 1) / NT;
    constexpr int AKP = BK + 4;                         // As row stride: rows stay 16-byte aligned
    static_assert((BM * QP) % NT == 0 && NT % QP == 0, "tile divisibility");

    __shared__ __align__(16) float As[BM][AKP];         // pixel-major: one STS.128 per produced float4
    __shared__ __align__(16) float Ws[BK][BN];

    const int tid = threadIdx.x;
    const int tx = tid % TXN, ty = tid / TXN;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int HoWo = p.Ho * p.Wo;

    // --- per-thread A-load assignment: element e = tid + i*NT ; kq = e % QP (same for every i) ; m = e / QP.
    // Consecutive lanes cover consecutive channel quads of one pixel, so a warp's 128-bit loads are contiguous.
    const int a_kq = tid % QP;
    const int a_mb = tid / QP;
    long long a_off[A_PER_THREAD];                      // element offset of the image (PW: of the pixel)
    int a_yx[A_PER_THREAD];                             // (iy0 << 16) | (ix0 & 0xffff); invalid pixel: a_off < 0
#pragma unroll
    for (int i = 0; i < A_PER_THREAD; i++) {
        const int gm = m0 + a_mb + i * (NT / QP);
        if (gm < p.M) {
            const int img = gm / HoWo;
            const int r = gm - img * HoWo;
            const int oy = r / p.Wo, ox = r - oy * p.Wo;
            a_off[i] = (long long)img * p.in_img_stride;
            if (MODE == CONV_PW) a_off[i] += ((long long)oy * p.W + ox) * p.Cs_in;
            a_yx[i] = ((oy * p.sh - p.pt) << 16) | ((ox * p.sw - p.pl) & 0xffff);
        } else {
            a_off[i] = -1;
            a_yx[i] = 0;
        }
    }

    float4 a_reg[A_PER_THREAD];
    float4 w_reg[W_PER_THREAD];

    auto load_chunk = [&](int k0) {
        const int k = k0 + a_kq * 4;
#pragma unroll
        for (int i = 0; i < A_PER_THREAD; i++) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a_off[i] >= 0 && k < p.K) {
                const int iy0 = a_yx[i] >> 16, ix0 = (int)(short)(a_yx[i] & 0xffff);
                if (MODE == CONV_PW) {
                    v = ldg4(p.in + a_off[i] + k);
                } else if (MODE == CONV_GATHER) {
                    const int tap = k / p.Cs_in, c = k - tap * p.Cs_in;
                    const int ky = tap / p.kw, kx = tap - ky * p.kw;
                    const int iy = iy0 + ky, ix = ix0 + kx;
                    if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
                        v = ldg4(p.in + a_off[i] + ((long long)iy * p.W + ix) * p.Cs_in + c);
                } else {   // CONV_DWPW: depthwise KS x KS over channels [k, k+4)
                    v = dw_window<KS == 0 ? 3 : KS>(p, p.in + a_off[i], iy0, ix0, k);
                    act4(v, p.act_mid, k);
                }
            }
            a_reg[i] = v;
        }
#pragma unroll
        for (int i = 0; i < W_PER_THREAD; i++) {
            const int e = tid + i * NT;
            const int nq = e % (BN / 4), kr = e / (BN / 4);
            const int gk = k0 + kr, gn = n0 + nq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < W_TOTAL && gk < p.K && gn < p.Ns) v = ldg4(p.w + (long long)gk * p.Ns + gn);
            w_reg[i] = v;
        }
    };

    auto store_chunk = [&]() {
#pragma unroll
        for (int i = 0; i < A_PER_THREAD; i++)
            *reinterpret_cast<float4 *>(&As[a_mb + i * (NT / QP)][a_kq * 4]) = a_reg[i];
#pragma unroll
        for (int i = 0; i < W_PER_THREAD; i++) {
            const int e = tid + i * NT;
            if (e < W_TOTAL) *reinterpret_cast<float4 *>(&Ws[e / (BN / 4)][(e % (BN / 4)) * 4]) = w_reg[i];
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

    load_chunk(0);
    for (int k0 = 0; k0 < p.K; k0 += BK) {
        store_chunk();
        __syncthreads();
        if (k0 + BK < p.K) load_chunk(k0 + BK);
        const int kmax = min(BK, p.K - k0);
#pragma unroll 2
        for (int kk = 0; kk < kmax; kk += 4) {
            float a4[TM][4], b4[4][TN];
#pragma unroll
            for (int i = 0; i < TM; i++) {
                const float4 v = *reinterpret_cast<const float4 *>(&As[ty * TM + i][kk]);
                a4[i][0] = v.x, a4[i][1] = v.y, a4[i][2] = v.z, a4[i][3] = v.w;
            }
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int j = 0; j < TN; j += 4) {
                    const float4 v = *reinterpret_cast<const float4 *>(&Ws[kk + q][tx * TN + j]);
                    b4[q][j] = v.x, b4[q][j + 1] = v.y, b4[q][j + 2] = v.z, b4[q][j + 3] = v.w;
                }
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int i = 0; i < TM; i++)
#pragma unroll
                    for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a4[i][q], b4[q][j], acc[i][j]);
        }
        __syncthreads();
    }

    // --- epilogue ------------------------------------------------------------------------------------
    const EpiDev &e = p.epi;
    const bool vec_ok = (p.out_pix_stride % 4 == 0) && (p.Nstore % 4 == 0) && (p.out_img_stride % 4 == 0);
    // residual of the whole micro-tile first: TM * TN / 4 independent 128-bit loads in flight instead of one
    // dependent load in front of every `v += residual` (the same stall ncu showed in the tcgen05 kernel's epilogue)
    const bool res_vec = e.res && (e.res_Cs % 4) == 0;
    float4 rpre[TM][TN / 4];
#pragma unroll
    for (int i = 0; i < TM; i++) {
        const int m = m0 + ty * TM + i;
        const int img = m / HoWo;
        const int r = m - img * HoWo;
        const int oy = r / p.Wo, ox = r - oy * p.Wo;
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
            const int n = n0 + tx * TN + j;
            rpre[i][j / 4] = (res_vec && m < p.M && n < p.Nstore) ? residual4_at(e, img, oy, ox, n) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
#pragma unroll
    for (int i = 0; i < TM; i++) {
        const int m = m0 + ty * TM + i;
        if (m >= p.M) continue;
        const int img = m / HoWo;
        const int r = m - img * HoWo;
        const int oy = r / p.Wo, ox = r - oy * p.Wo;
        float *orow = p.out + (long long)img * p.out_img_stride + (long long)r * p.out_pix_stride;
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
            const int n = n0 + tx * TN + j;
            if (n >= p.Nstore) continue;
            float v[4] = {acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]};
            if (n + 3 < p.Ns) {
                const float4 b = ldg4(e.bias + n);
                v[0] += b.x, v[1] += b.y, v[2] += b.z, v[3] += b.w;
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (n + q < p.Ns) v[q] += __ldg(e.bias + n + q);
            }
            act4(v, e.act1, n);
            if (e.res) {
                if ((e.res_Cs % 4) == 0) {
                    const float4 rr = rpre[i][j / 4];
                    v[0] += rr.x, v[1] += rr.y, v[2] += rr.z, v[3] += rr.w;
                } else {
#pragma unroll
                    for (int q = 0; q < 4; q++) v[q] += residual_at(e, img, oy, ox, n + q);
                }
            }
            act4(v, e.act2, n);
            if (vec_ok) {
                *reinterpret_cast<float4 *>(orow + n) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (n + q < p.Nstore) orow[n + q] = v[q];
            }
        }
    }
}

template <int BM, int BN, int TM, int TN, int BKT>
void launch_conv_cfg(const ConvDev &p, ConvMode mode, cudaStream_t s) {
    dim3 grid((p.M + BM - 1) / BM, (p.Ns + BN - 1) / BN);
    dim3 block((BM / TM) * (BN / TN));
    switch (mode) {
        case CONV_GATHER:
            ZB_KNAME("conv_gemm_kernel", BM, BN, TM, TN, BKT, CONV_GATHER, 0);
            conv_gemm_kernel<BM, BN, TM, TN, BKT, CONV_GATHER, 0><<<grid, block, 0, s>>>(p);
            break;
        case CONV_PW:
            ZB_KNAME("conv_gemm_kernel", BM, BN, TM, TN, BKT, CONV_PW, 0);
            conv_gemm_kernel<BM, BN, TM, TN, BKT, CONV_PW, 0><<<grid, block, 0, s>>>(p);
            break;
        case CONV_DWPW:
            if (p.kh == 3 && p.kw == 3) {
                ZB_KNAME("conv_gemm_kernel", BM, BN, TM, TN, BKT, CONV_DWPW, 3);
                conv_gemm_kernel<BM, BN, TM, TN, BKT, CONV_DWPW, 3><<<grid, block, 0, s>>>(p);
            } else if (p.kh == 5 && p.kw == 5) {
                ZB_KNAME("conv_gemm_kernel", BM, BN, TM, TN, BKT, CONV_DWPW, 5);
                conv_gemm_kernel<BM, BN, TM, TN, BKT, CONV_DWPW, 5><<<grid, block, 0, s>>>(p);
            }
            else throw std::runtime_error("unsupported op: fused depthwise kernel other than 3x3 / 5x5");
            break;
    }
}

template <int BM, int BN, int TM, int TN>
void launch_conv_tile(const ConvDev &p, ConvMode mode, cudaStream_t s) {
    if (p.K <= 16) launch_conv_cfg<BM, BN, TM, TN, 16>(p, mode, s);
    else launch_conv_cfg<BM, BN, TM, TN, 32>(p, mode, s);
}

// ------------------------------------------------------------------------------------------------
// Standalone depthwise conv: one thread per (output pixel, 4 channels).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dw_kernel(const ConvDev p) {
    const int cq_n = p.Cs_in / 4;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)p.M * cq_n) return;
    const int cq = (int)(idx % cq_n);
    const int m = (int)(idx / cq_n);
    const int HoWo = p.Ho * p.Wo;
    const int img = m / HoWo;
    const int r = m - img * HoWo;
    const int oy = r / p.Wo, ox = r - oy * p.Wo;
    const int c = cq * 4;
    const float *base = p.in + (long long)img * p.in_img_stride;
    const int iy0 = oy * p.sh - p.pt, ix0 = ox * p.sw - p.pl;
    float4 v = ldg4(p.epi.bias + c);
    for (int ky = 0; ky < p.kh; ky++) {
        const int iy = iy0 + ky;
        if (iy < 0 || iy >= p.H) continue;
        for (int kx = 0; kx < p.kw; kx++) {
            const int ix = ix0 + kx;
            if (ix < 0 || ix >= p.W) continue;
            const float4 x = ldg4(base + ((long long)iy * p.W + ix) * p.Cs_in + c);
            const float4 wv = ldg4(p.w + (ky * p.kw + kx) * p.Cs_in + c);
            v.x = fmaf(x.x, wv.x, v.x);
            v.y = fmaf(x.y, wv.y, v.y);
            v.z = fmaf(x.z, wv.z, v.z);
            v.w = fmaf(x.w, wv.w, v.w);
        }
    }
    const EpiDev &e = p.epi;
    float o[4] = {v.x, v.y, v.z, v.w};
    act4(o, e.act1, c);
    if (e.res) {
        const float4 rr = residual4_at(e, img, oy, ox, c);
        o[0] += rr.x, o[1] += rr.y, o[2] += rr.z, o[3] += rr.w;
    }
    act4(o, e.act2, c);
    float *orow = p.out + (long long)img * p.out_img_stride + (long long)r * p.out_pix_stride;
    *reinterpret_cast<float4 *>(orow + c) = make_float4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(256) maxpool2_kernel(const float *in, long long in_img_stride, int H, int W, int Cs,
                                                       float *out, long long out_img_stride, int n) {
    const int Ho = (H - 2) / 2 + 1, Wo = (W - 2) / 2 + 1, cq_n = Cs / 4;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n * Ho * Wo * cq_n) return;
    const int cq = (int)(idx % cq_n);
    long long m = idx / cq_n;
    const int ox = (int)(m % Wo);
    m /= Wo;
    const int oy = (int)(m % Ho);
    const int img = (int)(m / Ho);
    const float *p = in + (long long)img * in_img_stride + ((long long)(2 * oy) * W + 2 * ox) * Cs + cq * 4;
    const float4 a = ldg4(p), b = ldg4(p + Cs), c = ldg4(p + (long long)W * Cs), d = ldg4(p + (long long)W * Cs + Cs);
    float4 o = make_float4(fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y)),
                           fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z)), fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w)));
    *reinterpret_cast<float4 *>(out + (long long)img * out_img_stride + ((long long)oy * Wo + ox) * Cs + cq * 4) = o;
}

// ONNX Resize, mode=linear, coordinate_transformation_mode=half_pixel, scale 2 (palm FPN).
__global__ void __launch_bounds__(256) resize2x_kernel(const float *in, long long in_img_stride, int H, int W, int Cs,
                                                       float *out, long long out_img_stride, int n) {
    const int Ho = 2 * H, Wo = 2 * W, cq_n = Cs / 4;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n * Ho * Wo * cq_n) return;
    const int cq = (int)(idx % cq_n);
    long long m = idx / cq_n;
    const int ox = (int)(m % Wo);
    m /= Wo;
    const int oy = (int)(m % Ho);
    const int img = (int)(m / Ho);
    float sy = fmaxf((oy + 0.5f) * 0.5f - 0.5f, 0.0f), sx = fmaxf((ox + 0.5f) * 0.5f - 0.5f, 0.0f);
    int y0 = min((int)sy, H - 1), x0 = min((int)sx, W - 1);
    int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float fy = sy - (float)y0, fx = sx - (float)x0;
    const float *b = in + (long long)img * in_img_stride + cq * 4;
    const float4 v00 = ldg4(b + ((long long)y0 * W + x0) * Cs), v01 = ldg4(b + ((long long)y0 * W + x1) * Cs);
    const float4 v10 = ldg4(b + ((long long)y1 * W + x0) * Cs), v11 = ldg4(b + ((long long)y1 * W + x1) * Cs);
    auto lerp2 = [&](float a00, float a01, float a10, float a11) {
        const float top = a00 + (a01 - a00) * fx, bot = a10 + (a11 - a10) * fx;
        return top + (bot - top) * fy;
    };
    float4 o = make_float4(lerp2(v00.x, v01.x, v10.x, v11.x), lerp2(v00.y, v01.y, v10.y, v11.y),
                           lerp2(v00.z, v01.z, v10.z, v11.z), lerp2(v00.w, v01.w, v10.w, v11.w));
    *reinterpret_cast<float4 *>(out + (long long)img * out_img_stride + ((long long)oy * Wo + ox) * Cs + cq * 4) = o;
}

__global__ void __launch_bounds__(256) gap_kernel(const float *in, long long in_img_stride, int H, int W, int Cs,
                                                  float *out, long long out_img_stride, int n) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n * Cs) return;
    const int c = (int)(idx % Cs);
    const int img = (int)(idx / Cs);
    const float *b = in + (long long)img * in_img_stride + c;
    float acc = 0.f;
    const int hw = H * W;
    for (int i = 0; i < hw; i++) acc += __ldg(b + (long long)i * Cs);
    out[(long long)img * out_img_stride + c] = acc / (float)hw;
}

__global__ void __launch_bounds__(256) eltwise_kernel(const float *in, long long in_img_stride, int H, int W, int Cs,
                                                      float *out, long long out_img_stride, int out_pix_stride,
                                                      int Nstore, const EpiDev e, int n) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n * H * W * Cs) return;
    const int c = (int)(idx % Cs);
    long long m = idx / Cs;
    const int ox = (int)(m % W);
    m /= W;
    const int oy = (int)(m % H);
    const int img = (int)(m / H);
    if (c >= Nstore) return;
    float v = __ldg(in + (long long)img * in_img_stride + ((long long)oy * W + ox) * Cs + c);
    v = apply_act(v, e.act1, c);
    if (e.res) v += residual_at(e, img, oy, ox, c);
    v = apply_act(v, e.act2, c);
    out[(long long)img * out_img_stride + ((long long)oy * W + ox) * out_pix_stride + c] = v;
}

__global__ void __launch_bounds__(256) nchw_to_nhwc4_kernel(const float *in, int n, int H, int W, float *out,
                                                            long long out_img_stride, int f16) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long hw = (long long)H * W;
    if (idx >= (long long)n * hw) return;
    const int img = (int)(idx / hw);
    const long long pix = idx - (long long)img * hw;
    const float *b = in + (long long)img * 3 * hw + pix;
    float4 o = make_float4(__ldg(b), __ldg(b + hw), __ldg(b + 2 * hw), 0.f);
    if (f16) o.x = __half2float(__float2half_rn(o.x)), o.y = __half2float(__float2half_rn(o.y)), o.z = __half2float(__float2half_rn(o.z));
    *reinterpret_cast<float4 *>(out + (long long)img * out_img_stride + pix * 4) = o;
}

__global__ void __launch_bounds__(256) round_f16_kernel(float *data, long long count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) data[i] = __half2float(__float2half_rn(data[i]));
}

inline unsigned blocks_for(long long total, int bs) { return (unsigned)((total + bs - 1) / bs); }

}  // namespace

void launch_conv(const ConvDev &p, ConvMode mode, cudaStream_t s) {
    g_launch_count++;
    if (mode == CONV_DWPW && launch_dwpw_thin(p, s)) return;
    if (mode == CONV_PW && launch_pw_thin(p, s)) return;
    if (p.Ns <= 16) launch_conv_tile<256, 16, 4, 4>(p, mode, s);
    else if (p.Ns <= 32) launch_conv_tile<128, 32, 4, 4>(p, mode, s);
    else if (p.Ns <= 64) launch_conv_tile<128, 64, 8, 4>(p, mode, s);
    else launch_conv_tile<64, 128, 4, 8>(p, mode, s);
}

void launch_dw(const ConvDev &p, cudaStream_t s) {
    g_launch_count++;
    long long total = (long long)p.M * (p.Cs_in / 4);
    ZB_KNAME("dw_kernel");
    dw_kernel<<<blocks_for(total, 256), 256, 0, s>>>(p);
}

void launch_maxpool2(const float *in, long long in_img_stride, int H, int W, int Cs, float *out,
                     long long out_img_stride, int n, cudaStream_t s) {
    g_launch_count++;
    long long total = (long long)n * ((H - 2) / 2 + 1) * ((W - 2) / 2 + 1) * (Cs / 4);
    ZB_KNAME("maxpool2_kernel");
    maxpool2_kernel<<<blocks_for(total, 256), 256, 0, s>>>(in, in_img_stride, H, W, Cs, out, out_img_stride, n);
}

void launch_resize2x(const float *in, long long in_img_stride, int H, int W, int Cs, float *out,
                     long long out_img_stride, int n, cudaStream_t s) {
    g_launch_count++;
    long long total = (long long)n * 4 * H * W * (Cs / 4);
    ZB_KNAME("resize2x_kernel");
    resize2x_kernel<<<blocks_for(total, 256), 256, 0, s>>>(in, in_img_stride, H, W, Cs, out, out_img_stride, n);
}

void launch_gap(const float *in, long long in_img_stride, int H, int W, int Cs, float *out, long long out_img_stride,
                int n, cudaStream_t s) {
    g_launch_count++;
    ZB_KNAME("gap_kernel");
    gap_kernel<<<blocks_for((long long)n * Cs, 256), 256, 0, s>>>(in, in_img_stride, H, W, Cs, out, out_img_stride, n);
}

void launch_eltwise(const float *in, long long in_img_stride, int H, int W, int Cs, float *out,
                    long long out_img_stride, int out_pix_stride, int Nstore, const EpiDev &epi, int n,
                    cudaStream_t s) {
    g_launch_count++;
    long long total = (long long)n * H * W * Cs;
    ZB_KNAME("eltwise_kernel");
    eltwise_kernel<<<blocks_for(total, 256), 256, 0, s>>>(in, in_img_stride, H, W, Cs, out, out_img_stride,
                                                          out_pix_stride, Nstore, epi, n);
}

void launch_nchw_to_nhwc4(const float *in_nchw, int n, int H, int W, float *out, long long out_img_stride,
                          cudaStream_t s, int round_f16) {
    g_launch_count++;
    ZB_KNAME("nchw_to_nhwc4_kernel");
    nchw_to_nhwc4_kernel<<<blocks_for((long long)n * H * W, 256), 256, 0, s>>>(in_nchw, n, H, W, out, out_img_stride, round_f16);
}

void launch_round_f16(float *data, long long count, cudaStream_t s) {
    g_launch_count++;
    ZB_KNAME("round_f16_kernel");
    round_f16_kernel<<<blocks_for(count, 256), 256, 0, s>>>(data, count);
}

}  // namespace zb
