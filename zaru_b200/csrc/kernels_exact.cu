// sm_100a kernels whose arithmetic must reproduce the reference's scalar f32 code bit for bit:
// image->tensor sampling, SSD decode + NMS + remap, RoI/view algebra, landmark unpacking.
// THIS FILE IS COMPILED WITH -fmad=false: every mul/add/sub/div rounds separately, like Rust.
//
//   sample_kernel ....... Cnn image_map closure + sample + ViewData::image_coord + ColorMapper::map
//                         (crates/zaru/src/nn/mod.rs:54-73, :156-166; image/mod.rs:224-247)
//   decode_nms_kernel ... extract_outputs / extract_detection (face/detection.rs:96-157,
//                         hand/detection.rs:108-179), Anchors::calculate (detection/ssd.rs:96-119),
//                         NonMaxSuppression::process (detection/nms.rs:59-145), Rect::iou
//                         (zaru-image/src/rect.rs:193-214), remap (detection.rs:245-267)
//   face_roi_kernel ..... examples/facemesh.rs:49-54 + LandmarkTracker::track (landmark.rs:465-466)
//                         + Estimator::estimate_impl view fitting (landmark.rs:320-323)
//   landmarks_kernel .... Network::extract impls + Estimator remap (landmark.rs:336-345) + tracker
//                         transform_out (landmark.rs:482-486)
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math_constants.h>

#include <cstdlib>
#include <type_traits>

#include "geom.h"
#include "kernels.h"
#include "tc_common.cuh"

namespace zb {
namespace {

// num.rs:6-8 `1.0 / (1.0 + (-v).exp())`.  Rust's f32::exp is glibc expf (< 0.51 ulp, i.e. correctly rounded
// in practice); CUDA's expf is only 2-ulp accurate, so evaluate exp in f64 and round once.
__device__ __forceinline__ float sigmoid_ref(float v) { return 1.0f / (1.0f + (float)exp((double)(-v))); }

// Rust `as u32` on a rounded f32 kept in float form (saturating, NaN -> 0), then `as f32`.
__device__ __forceinline__ float sat_u32_as_f32(float v) {
    if (!(v > 0.0f)) return 0.0f;                 // negatives and NaN -> 0
    if (v >= 4294967296.0f) return 4294967296.0f; // u32::MAX as f32
    return v;
}

// ------------------------------------------------------------------------------------------------
// One sampled, colour-mapped pixel of the network input tensor: (r, g, b, 0).
//   sample(): x = (u * view_w).round() as u32 with u = x as f32 / w as f32          (nn/mod.rs:54-58)
//   ViewData::image_coord: transform_out(x+0.5, y+0.5), round(p-0.5), bounds          (image/mod.rs:224-241)
//   ColorMapper::map: col as f32 * ((end - start) / 255.0) + start                   (nn/mod.rs:156-166)
// Address of the source texel for tensor pixel (x, y), or nullptr when it reads Color::NONE.
__device__ __forceinline__ const unsigned *sample_address(const FramesDev &f, const ViewDev &v, int x, int y, int out_w,
                                                          int out_h) {
    if (!v.valid) return nullptr;
    const int xs = v.flip_x ? (out_w - 1 - x) : x;
    const float u = (float)xs / (float)out_w;
    const float vv = (float)y / (float)out_h;
    const float sx = sat_u32_as_f32(roundf(u * v.w));
    const float sy = sat_u32_as_f32(roundf(vv * v.h));
    RRectF rr;
    rr.r.cx = v.cx, rr.r.cy = v.cy, rr.r.w = v.w, rr.r.h = v.h;
    rr.c = v.cosr, rr.s = v.sinr, rr.rad = 0.f;
    float px, py;
    transform_out(rr, sx + 0.5f, sy + 0.5f, px, py);
    const float fx = roundf(px - 0.5f), fy = roundf(py - 0.5f);
    if (fx < 0.0f || fy < 0.0f || ceilf(fx) >= 4294967296.0f || ceilf(fy) >= 4294967296.0f) return nullptr;
    const unsigned ix = (fx == fx) ? (unsigned)fx : 0u;   // `x.round() as u32`: NaN -> 0
    const unsigned iy = (fy == fy) ? (unsigned)fy : 0u;
    if (ix >= (unsigned)f.width || iy >= (unsigned)f.height) return nullptr;
    return reinterpret_cast<const unsigned *>(f.base + (long long)v.frame * f.frame_stride + (long long)iy * f.row_stride +
                                              (long long)ix * 4);
}
// f16 != 0: the network's input is FLOAT16 - NeuralNetwork::estimate rounds every value with half::f16::from_f32
// (round to nearest even) before inference (nn/mod.rs:487-492); the device keeps the rounded value in an f32.
__device__ __forceinline__ float4 color_map(unsigned rgba, float lo, float adjust, int f16 = 0) {
    float4 c = make_float4((float)(rgba & 0xFFu) * adjust + lo, (float)((rgba >> 8) & 0xFFu) * adjust + lo,
                           (float)((rgba >> 16) & 0xFFu) * adjust + lo, 0.0f);
    if (f16) c.x = __half2float(__float2half_rn(c.x)), c.y = __half2float(__float2half_rn(c.y)), c.z = __half2float(__float2half_rn(c.z));
    return c;
}

__device__ __forceinline__ float4 sample_pixel(const FramesDev &f, const ViewDev &v, int x, int y, int out_w, int out_h,
                                               float lo, float adjust, int f16 = 0) {
    unsigned rgba = 0u;   // Color::NONE
    if (v.valid) {
        const int xs = v.flip_x ? (out_w - 1 - x) : x;
        const float u = (float)xs / (float)out_w;
        const float vv = (float)y / (float)out_h;
        const float sx = sat_u32_as_f32(roundf(u * v.w));
        const float sy = sat_u32_as_f32(roundf(vv * v.h));
        RRectF rr;
        rr.r.cx = v.cx, rr.r.cy = v.cy, rr.r.w = v.w, rr.r.h = v.h;
        rr.c = v.cosr, rr.s = v.sinr, rr.rad = 0.f;
        float px, py;
        transform_out(rr, sx + 0.5f, sy + 0.5f, px, py);
        const float fx = roundf(px - 0.5f), fy = roundf(py - 0.5f);
        const bool reject = fx < 0.0f || fy < 0.0f || ceilf(fx) >= 4294967296.0f || ceilf(fy) >= 4294967296.0f;
        if (!reject) {
            // `x.round() as u32`: NaN -> 0
            const unsigned ix = (fx == fx) ? (unsigned)fx : 0u;
            const unsigned iy = (fy == fy) ? (unsigned)fy : 0u;
            if (ix < (unsigned)f.width && iy < (unsigned)f.height) {
                const uint8_t *p = f.base + (long long)v.frame * f.frame_stride + (long long)iy * f.row_stride +
                                   (long long)ix * 4;
                rgba = __ldg(reinterpret_cast<const unsigned *>(p));
            }
        }
    }
    return color_map(rgba, lo, adjust, f16);
}

// Persistent small grid: each thread keeps 4 independent host-memory reads in flight per iteration.
__global__ void __launch_bounds__(256) gather_texels_kernel(const FramesDev f, const ViewDev *__restrict__ views, int n,
                                                            int out_w, int out_h, unsigned *__restrict__ out) {
    const long long per = (long long)out_w * out_h, total = per * n;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
        const unsigned *addr[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = i0 + u * stride;
            addr[u] = nullptr;
            if (i < total) {
                const int img = (int)(i / per);
                const int r = (int)(i - (long long)img * per);
                const int y = r / out_w, x = r - y * out_w;
                addr[u] = sample_address(f, views[img], x, y, out_w, out_h);
            }
        }
        unsigned v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = addr[u] ? __ldg(addr[u]) : 0u;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = i0 + u * stride;
            if (i < total) out[i] = v[u];
        }
    }
}

// ImageView::get for every pixel of the view (no resampling: view pixel (x, y) -> image_coord -> texel or NONE)
__global__ void __launch_bounds__(128) view_to_image_kernel(const FramesDev f, const ViewDev *__restrict__ views, int out_w,
                                                            int out_h, unsigned *__restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, img = blockIdx.z;
    if (x >= out_w) return;
    const ViewDev v = views[img];
    unsigned rgba = 0u;                                  // Color::NONE
    if (v.valid) {
        RRectF rr;
        rr.r.cx = v.cx, rr.r.cy = v.cy, rr.r.w = v.w, rr.r.h = v.h;
        rr.c = v.cosr, rr.s = v.sinr, rr.rad = 0.f;
        float px, py;
        transform_out(rr, (float)x + 0.5f, (float)y + 0.5f, px, py);
        const float fx = roundf(px - 0.5f), fy = roundf(py - 0.5f);
        const bool reject = fx < 0.0f || fy < 0.0f || ceilf(fx) >= 4294967296.0f || ceilf(fy) >= 4294967296.0f;
        if (!reject) {
            const unsigned ix = (fx == fx) ? (unsigned)fx : 0u, iy = (fy == fy) ? (unsigned)fy : 0u;
            if (ix < (unsigned)f.width && iy < (unsigned)f.height)
                rgba = __ldg(reinterpret_cast<const unsigned *>(f.base + (long long)v.frame * f.frame_stride +
                                                                (long long)iy * f.row_stride + (long long)ix * 4));
        }
    }
    out[((long long)img * out_h + y) * out_w + x] = rgba;
}

__global__ void __launch_bounds__(256) frames_clear_kernel(uint8_t *base, long long frame_stride, long long row_stride, int width,
                                                           int height, int first, unsigned rgba) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= width) return;
    *reinterpret_cast<unsigned *>(base + (long long)(first + blockIdx.z) * frame_stride + (long long)blockIdx.y * row_stride +
                                  (long long)x * 4) = rgba;
    (void)height;
}

__global__ void __launch_bounds__(256) sample_kernel(const FramesDev f, const ViewDev *__restrict__ views, int out_w,
                                                     int out_h, float lo, float hi, int layout,
                                                     float *__restrict__ out, long long out_img_stride, int f16) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int img = blockIdx.z;
    if (x >= out_w) return;
    const ViewDev v = views[img];
    const float adjust = (hi - lo) / 255.0f;
    const float4 c = sample_pixel(f, v, x, y, out_w, out_h, lo, adjust, f16);
    float *o = out + (long long)img * out_img_stride;
    const long long pix = (long long)y * out_w + x;
    if (layout == SAMPLE_NHWC4) {
        *reinterpret_cast<float4 *>(o + pix * 4) = c;
    } else if (layout == SAMPLE_NHWC3) {
        o[pix * 3 + 0] = c.x, o[pix * 3 + 1] = c.y, o[pix * 3 + 2] = c.z;
    } else {
        const long long hw = (long long)out_w * out_h;
        o[pix] = c.x, o[hw + pix] = c.y, o[2 * hw + pix] = c.z;
    }
}

// ------------------------------------------------------------------------------------------------
// Fused image->tensor sampling + stem convolution (KS x KS, stride 2, 3 -> N <= 64 channels) + bias + act.
// The sampled network-input tile a CTA needs is produced straight into shared memory (bit-exact sampler
// above), so the [n,h,w,3] input tensor never exists in HBM; with `views == nullptr` the tile is read from
// an NHWC4 tensor instead (NeuralNetwork::estimate on caller tensors).  One thread = one output pixel, all
// N accumulators in registers, weights read from shared memory as warp broadcasts.  The convolution uses
// explicit fmaf (this translation unit is built with -fmad=false for the sampler's sake).
// ------------------------------------------------------------------------------------------------
// For an UNROTATED view (cos == 1, sin == 0) the source coordinate of tensor pixel (x, y) is separable:
//   transform_out gives px = ((0 + 1*dx) + (-0)*dy + hx) + tlx = (dx + hx) + tlx  exactly (adding a signed zero
//   never changes dx, and 0 + dx is never -0), and likewise py depends on y only.
// So one table entry per tile column / row replaces ~200 instructions of f32 division, rounding and bounds
// checks per sampled texel.  Returns the source column (row) or -1 when the texel reads Color::NONE.
__device__ __forceinline__ int sample_axis(float view_pos_center, float view_extent, int t, int out_n, int flip, int limit) {
    // view_pos_center / view_extent: RotatedRect centre and size along this axis
    const int ts = flip ? (out_n - 1 - t) : t;
    const float u = (float)ts / (float)out_n;
    const float sc = sat_u32_as_f32(roundf(u * view_extent));
    const float h = view_extent * 0.5f;
    const float d = (sc + 0.5f) - h;
    const float tl = view_pos_center - view_extent * 0.5f;
    const float pp = ((0.0f + d) + h) + tl;           // same operation order as transform_out with c = 1, s = 0
    const float fr = roundf(pp - 0.5f);
    if (fr < 0.0f || ceilf(fr) >= 4294967296.0f) return -1;
    const unsigned i = (fr == fr) ? (unsigned)fr : 0u;
    return i < (unsigned)limit ? (int)i : -1;
}

template <int KS, int NP, int PPT>
__global__ void __launch_bounds__(256) stem_kernel(const FramesDev f, const ViewDev *__restrict__ views, float lo, float hi,
                                                   const ConvDev p, int tiles_x, int tiles_y, int NSP, int f16) {
    constexpr int TW = 32, TH = 8 * PPT, NT = 256;               // PPT output pixels per thread (rows ty, ty+8)
    constexpr int IW = (TW - 1) * 2 + KS, IH = (TH - 1) * 2 + KS;
    extern __shared__ __align__(16) float smem[];
    float4 *s_in = reinterpret_cast<float4 *>(smem);            // [IH][IW] (r,g,b,0)
    float *s_w = smem + IH * IW * 4;                           // [KS*KS*4][NSP]  (row 4*tap+ci)
    float *s_b = s_w + KS * KS * 4 * NSP;                      // [NSP]
    float *s_sl = s_b + NSP;                                   // [NSP]
    int *s_col = reinterpret_cast<int *>(s_sl + NSP);          // [IW] source column per tile column (-1: NONE)
    int *s_row = s_col + IW;                                   // [IH]

    const int tid = threadIdx.x;
    int b = blockIdx.x;
    const int tile_x = b % tiles_x;
    b /= tiles_x;
    const int tile_y = b % tiles_y;
    const int img = b / tiles_y;
    const int oy0 = tile_y * TH, ox0 = tile_x * TW;
    const int iy_org = oy0 * 2 - p.pt, ix_org = ox0 * 2 - p.pl;

    if (views) {
        const ViewDev v = views[img];
        const float adjust = (hi - lo) / 255.0f;
        const bool separable = v.valid && v.cosr == 1.0f && v.sinr == 0.0f;
        if (separable) {
            // tile column / row -> source column / row, -2 marks conv padding (outside the tensor)
            for (int e = tid; e < IW + IH; e += NT) {
                if (e < IW) {
                    const int ix = ix_org + e;
                    s_col[e] = (ix >= 0 && ix < p.W) ? sample_axis(v.cx, v.w, ix, p.W, v.flip_x, f.width) : -2;
                } else {
                    const int iy = iy_org + (e - IW);
                    s_row[e - IW] = (iy >= 0 && iy < p.H) ? sample_axis(v.cy, v.h, iy, p.H, 0, f.height) : -2;
                }
            }
            __syncthreads();
            const uint8_t *fbase = f.base + (long long)v.frame * f.frame_stride;
            for (int e0 = tid; e0 < IH * IW; e0 += 4 * NT) {
                unsigned rgba[4];
                int kind[4];                                   // 0: conv padding, 1: Color::NONE, 2: texel
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int e = e0 + u * NT;
                    rgba[u] = 0u;
                    kind[u] = 0;
                    if (e < IH * IW) {
                        const int ty = e / IW, tx = e - ty * IW;
                        const int sc = s_col[tx], sr = s_row[ty];
                        if (sc != -2 && sr != -2) {
                            kind[u] = 1;
                            if (sc >= 0 && sr >= 0) {
                                kind[u] = 2;
                                rgba[u] = __ldg(reinterpret_cast<const unsigned *>(fbase + (long long)sr * f.row_stride + (long long)sc * 4));
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int e = e0 + u * NT;
                    if (e < IH * IW) s_in[e] = kind[u] ? color_map(rgba[u], lo, adjust, f16) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        } else {
            // general (rotated) view: per-texel coordinate math, batches of 4 texels per thread
            for (int e0 = tid; e0 < IH * IW; e0 += 4 * NT) {
                const unsigned *addr[4];
                bool inside[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int e = e0 + u * NT;
                    const int ty = e / IW, tx = e - ty * IW;
                    const int iy = iy_org + ty, ix = ix_org + tx;
                    inside[u] = e < IH * IW && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                    addr[u] = inside[u] ? sample_address(f, v, ix, iy, p.W, p.H) : nullptr;
                }
                unsigned rgba[4];
#pragma unroll
                for (int u = 0; u < 4; u++) rgba[u] = addr[u] ? __ldg(addr[u]) : 0u;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int e = e0 + u * NT;
                    if (e < IH * IW)   // conv zero padding outside the tensor (NOT the letterbox colour)
                        s_in[e] = inside[u] ? color_map(rgba[u], lo, adjust, f16) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    } else {
        const float *in_img = p.in + (long long)img * p.in_img_stride;
        for (int e = tid; e < IH * IW; e += NT) {
            const int ty = e / IW, tx = e - ty * IW;
            const int iy = iy_org + ty, ix = ix_org + tx;
            float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
            if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
                c = __ldg(reinterpret_cast<const float4 *>(in_img + ((long long)iy * p.W + ix) * 4));
            s_in[e] = c;
        }
    }
    for (int e = tid; e < KS * KS * 4 * NSP / 4; e += NT) {
        const int k = e / (NSP / 4), nq = e - k * (NSP / 4);
        float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nq * 4 < p.Ns) wv = __ldg(reinterpret_cast<const float4 *>(p.w + (long long)k * p.Ns + nq * 4));
        reinterpret_cast<float4 *>(s_w)[e] = wv;
    }
    for (int e = tid; e < NSP; e += NT) {
        s_b[e] = e < p.Ns ? __ldg(p.epi.bias + e) : 0.f;
        s_sl[e] = (p.epi.act1.kind == ACT_PRELU && e < p.Ns) ? __ldg(p.epi.act1.slope + e) : 0.f;
    }
    __syncthreads();

    const int tx = tid % TW, ty = tid / TW;
    const int ox = ox0 + tx;
    if (ox >= p.Wo) return;
    const int act = p.epi.act1.kind;
    for (int n0 = 0; n0 < p.Nstore; n0 += NP) {
        // accumulators as float2 pairs: packed FFMA2 (two fused multiply-adds per issued instruction, each lane
        // rounds exactly like fmaf) halves the FMA issue slots of this FMA-bound kernel
        float2 acc[PPT][NP / 2];
#pragma unroll
        for (int i = 0; i < PPT; i++)
#pragma unroll
            for (int j = 0; j < NP; j += 2) acc[i][j / 2] = make_float2(s_b[n0 + j], s_b[n0 + j + 1]);
#pragma unroll
        for (int t = 0; t < KS * KS; t++) {
            float xs[PPT][3];
#pragma unroll
            for (int i = 0; i < PPT; i++) {
                const float4 x = s_in[((ty + 8 * i) * 2 + t / KS) * IW + tx * 2 + (t % KS)];
                xs[i][0] = x.x, xs[i][1] = x.y, xs[i][2] = x.z;
            }
#pragma unroll
            for (int ci = 0; ci < 3; ci++) {
#pragma unroll
                for (int j = 0; j < NP; j += 4) {
                    const float4 wv = *reinterpret_cast<const float4 *>(s_w + (t * 4 + ci) * NSP + n0 + j);
#pragma unroll
                    for (int i = 0; i < PPT; i++) {       // each weight quad feeds PPT pixels
                        const float2 aa = make_float2(xs[i][ci], xs[i][ci]);
                        acc[i][j / 2] = __ffma2_rn(aa, make_float2(wv.x, wv.y), acc[i][j / 2]);
                        acc[i][j / 2 + 1] = __ffma2_rn(aa, make_float2(wv.z, wv.w), acc[i][j / 2 + 1]);
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < PPT; i++) {
            const int oy = oy0 + ty + 8 * i;
            if (oy >= p.Ho) continue;
            float *orow = p.out + (long long)img * p.out_img_stride + ((long long)oy * p.Wo + ox) * p.out_pix_stride;
#pragma unroll
            for (int j = 0; j < NP; j += 4) {
                const int n = n0 + j;
                if (n >= p.Nstore) break;
                float v[4] = {acc[i][j / 2].x, acc[i][j / 2].y, acc[i][j / 2 + 1].x, acc[i][j / 2 + 1].y};
                if (act == ACT_RELU) {
#pragma unroll
                    for (int q = 0; q < 4; q++) v[q] = fmaxf(v[q], 0.0f);
                } else if (act == ACT_PRELU) {
#pragma unroll
                    for (int q = 0; q < 4; q++) v[q] = v[q] < 0.0f ? v[q] * s_sl[n + q] : v[q];
                } else if (act == ACT_CLIP) {
#pragma unroll
                    for (int q = 0; q < 4; q++) v[q] = fminf(fmaxf(v[q], p.epi.act1.lo), p.epi.act1.hi);
                }
                *reinterpret_cast<float4 *>(orow + n) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    }
}

template <int KS, int NP, int PPT>
bool launch_stem_cfg(const FramesDev &f, const ViewDev *views, float lo, float hi, const ConvDev &p, int f16, cudaStream_t s) {
    constexpr int TW = 32, TH = 8 * PPT;
    constexpr int IW = (TW - 1) * 2 + KS, IH = (TH - 1) * 2 + KS;
    const int NSP = (p.Ns + NP - 1) / NP * NP;
    const size_t smem = sizeof(float) * ((size_t)IH * IW * 4 + (size_t)KS * KS * 4 * NSP + 2 * NSP + IW + IH);
    auto kern = stem_kernel<KS, NP, PPT>;
    static SmemOptIn opt_in;
    if (!opt_in.ensure(kern, smem)) return false;
    const int tiles_x = (p.Wo + TW - 1) / TW, tiles_y = (p.Ho + TH - 1) / TH;
    const int images = p.M / (p.Ho * p.Wo);
    ZB_KNAME("stem_kernel", KS, NP, PPT);
    kern<<<(unsigned)(tiles_x * tiles_y * images), 256, smem, s>>>(f, views, lo, hi, p, tiles_x, tiles_y, NSP, f16);
    return true;
}

// ------------------------------------------------------------------------------------------------
// The same stem on the tensor core ("stem_mma"): tcgen05.mma kind::f16 with the SAMPLED TEXEL TILE ITSELF as the A operand.
//
//   * An 8-bit texel value is exact in FP16, so the tile is stored once as (r, g, b, inside-the-tensor) FP16 quads - 8 bytes per
//     texel, no hi / lo split of the activation side - and the affine ColorMapper is folded into the weights:
//         sum_taps (a x + b) w  =  sum_taps x (a w)  +  sum_taps inside * (b (w_r + w_g + w_b))
//     (a texel that reads Color::NONE is x = 0 WITH inside = 1: it maps to `b`; a tap outside the tensor is conv padding: all 0).
//   * No im2col: in the K-major SWIZZLE_NONE layout the rows of a core matrix are 16 bytes apart = 2 texels = one output pixel of
//     a stride-2 conv, so with a leading-dimension offset of 16 bytes a descriptor pointing at texel (2 oy + ky, 2 ox) reads, for
//     the 8 x 16 output pixels of an MMA tile, the 4 consecutive taps (ky, 0..3) of every pixel as one K = 16 step - overlapping
//     rows of the same tile, nothing is copied.  A tap row costs one (3x3) or two (5x5; slots 5..7 carry zero weights) MMAs.
//   * Weights: f32 -> folded, scaled by a power of two into FP16's normal range, split into FP16 hi + lo (22 bits) by every CTA
//     while its texel gather is in flight; two MMA passes (lo, hi) accumulate in FP32 in TMEM.
//   * One CTA = 32 x 16 output pixels = four M = 128 accumulators; one barrier between the gather and the MMAs, one mbarrier
//     wait, and an epilogue that needs no CTA barrier (each warp stages its own 32 pixels and writes them out coalesced).
//     Three CTAs per SM overlap each other's gather / MMA / epilogue phases.
// 5x5 -> 24 (BlazeFace): 1800 FMAs per output pixel on the FP32 pipe become 20 MMAs per 128 pixels.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                       // c_format = F32; a_format = b_format = 0 (F16), both K-major
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// (r, g, b, inside) as four FP16 values: bytes -> 0x6400 | byte = 1024 + byte, minus 1024 (exact)
__device__ __forceinline__ uint2 texel_f16(unsigned rgba, bool inside) {
    unsigned rg = __byte_perm(rgba, 0x64006400u, 0x5150);   // (0x6400 | r) | (0x6400 | g) << 16
    unsigned bx = __byte_perm(rgba, 0x64006400u, 0x5452);   // (0x6400 | b) | 0x6400 << 16
    const __half2 k = __floats2half2_rn(1024.0f, 1024.0f);
    __half2 h0 = __hsub2(*reinterpret_cast<__half2 *>(&rg), k), h1 = __hsub2(*reinterpret_cast<__half2 *>(&bx), k);
    uint2 o;
    o.x = *reinterpret_cast<unsigned *>(&h0);
    o.y = (*reinterpret_cast<unsigned *>(&h1) & 0xffffu) | (inside ? 0x3c000000u : 0u);
    return inside ? o : make_uint2(0u, 0u);
}

// FLOAT16-input networks: NeuralNetwork::estimate rounds the mapped value to f16 (color_map, f16 = 1) - that value IS an FP16
// number, so it goes into the tile as it is: (f16(a r + b), f16(a g + b), f16(a b + b), 0); conv padding stays all zero.
__device__ __forceinline__ uint2 texel_f16_mapped(unsigned rgba, bool inside, float lo, float adjust) {
    if (!inside) return make_uint2(0u, 0u);
    const float4 c = color_map(rgba, lo, adjust, 1);
    const __half2 h0 = __floats2half2_rn(c.x, c.y), h1 = __floats2half2_rn(c.z, 0.0f);
    uint2 o;
    o.x = *reinterpret_cast<const unsigned *>(&h0);
    o.y = *reinterpret_cast<const unsigned *>(&h1);
    return o;
}

template <int KS, int N>
struct StemMma {
    static constexpr int NACC = N <= 32 ? 4 : 2;              // M = 128 accumulators (8 x 16 output pixels each) per CTA: 128 TMEM columns
    static constexpr int NLD = N <= 32 ? N : 32;              // accumulator columns a warp reads, finishes and stages at a time
    static constexpr int TW = 8 * NACC, TH = 16, NT = 256;
    static constexpr int KSTEPS = KS == 5 ? 2 : 1;            // K = 16 steps per tap row (4 taps each)
    static constexpr int KX = 4 * KSTEPS;                     // tap slots per row (zero weights beyond KS)
    static constexpr int IW = (TW - 1) * 2 + KS, IH = (TH - 1) * 2 + KS;
    static constexpr int P = ((TW - 1) * 2 + KX + 1) & ~1;    // texel pitch: the last pixel's padded tap slots stay inside the row
    static constexpr int TILE_BYTES = IH * P * 8;
    static constexpr int PART_BYTES = KS * (KX / 2) * N * 16; // [ky][k8][n][8 halves]
    static constexpr int STAGE_STRIDE = NLD + 4;              // floats; conflict-free row-per-lane 128-bit stores
    static constexpr int STAGE_BYTES = 8 * 32 * STAGE_STRIDE * 4;
    static constexpr int MAIN_BYTES = (TILE_BYTES + 2 * PART_BYTES > STAGE_BYTES ? TILE_BYTES + 2 * PART_BYTES : STAGE_BYTES);
    static constexpr int SMEM_BYTES = MAIN_BYTES + (2 * N + IW + IH) * 4;
    static constexpr int TCOLS = NACC * N < 32 ? 32 : NACC * N;
};

template <int KS, int N, int MINB>
__global__ void __launch_bounds__(256, MINB) stem_mma_kernel(const FramesDev f, const ViewDev *__restrict__ views, float lo, float hi,
                                                          const ConvDev p, int tiles_x, int tiles_y, int f16) {
    using G = StemMma<KS, N>;
    constexpr int TW = G::TW, TH = G::TH, NT = G::NT, KX = G::KX, IW = G::IW, IH = G::IH, P = G::P;
    extern __shared__ __align__(128) unsigned char stem_mma_smem[];
    uint2 *s_tile = reinterpret_cast<uint2 *>(stem_mma_smem);                                  // [IH][P] FP16 quads
    unsigned char *s_B = stem_mma_smem + G::TILE_BYTES;                                       // [part][ky][k8][N][8 halves]
    float *s_stage = reinterpret_cast<float *>(stem_mma_smem);                                // epilogue: aliases tile + weights
    float *s_bias = reinterpret_cast<float *>(stem_mma_smem + G::MAIN_BYTES);                 // [N]
    float *s_sl = s_bias + N;                                                            // [N]
    int *s_col = reinterpret_cast<int *>(s_sl + N);                                      // [IW]
    int *s_row = s_col + IW;                                                             // [IH]
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_red[8];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int b = blockIdx.x;
    const int tile_x = b % tiles_x;
    b /= tiles_x;
    const int tile_y = b % tiles_y;
    const int img = b / tiles_y;
    const int oy0 = tile_y * TH, ox0 = tile_x * TW;
    const int iy_org = oy0 * 2 - p.pt, ix_org = ox0 * 2 - p.pl;
    pdl_wait();                                                       // the views may come from the kernel in front (RoI / compaction)
    const ViewDev v = views[img];
    const float adjust = (hi - lo) / 255.0f;
    const bool separable = v.valid && v.cosr == 1.0f && v.sinr == 0.0f;

    if (warp == 0) tc::tmem_alloc(&tmem_slot, G::TCOLS);
    const int jmax = min(G::NACC, (p.Wo - ox0 + 7) >> 3);            // MMA tiles (8 output columns each) inside the map
    if (tid == 0) tc::mbar_init(&mbar, jmax);                         // one commit per issuing thread
    if (separable) {
        for (int e = tid; e < IW + IH; e += NT) {
            if (e < IW) {
                const int ix = ix_org + e;
                s_col[e] = (ix >= 0 && ix < p.W) ? sample_axis(v.cx, v.w, ix, p.W, v.flip_x, f.width) : -2;
            } else {
                const int iy = iy_org + (e - IW);
                s_row[e - IW] = (iy >= 0 && iy < p.H) ? sample_axis(v.cy, v.h, iy, p.H, 0, f.height) : -2;
            }
        }
    }
    for (int e = tid; e < N; e += NT) {
        s_bias[e] = e < p.Ns ? __ldg(p.epi.bias + e) : 0.f;
        s_sl[e] = (p.epi.act1.kind == ACT_PRELU && e < p.Ns) ? __ldg(p.epi.act1.slope + e) : 0.f;
    }
    // folded weights of this thread's (ky, tap slot, n) items: (a w_r, a w_g, a w_b, b (w_r + w_g + w_b)); zero beyond KS / Ns
    constexpr int ITEMS = KS * KX * N, PER = (ITEMS + NT - 1) / NT;
    float4 wv[PER];
    float wmax = 0.0f;
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const int e = tid + i * NT;
        wv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < ITEMS) {
            const int n = e % N, t = e / N, kx = t % KX, ky = t / KX;
            if (kx < KS && n < p.Ns) {
                const float *wp = p.w + (long long)((ky * KS + kx) * 4) * p.Ns + n;
                const float wr = __ldg(wp), wg = __ldg(wp + p.Ns), wb = __ldg(wp + 2 * p.Ns);
                // FLOAT16 input (FaceMeshV2): the tile holds f16(a x + b) itself, the weights stay as they are
                wv[i] = f16 ? make_float4(wr, wg, wb, 0.0f) : make_float4(adjust * wr, adjust * wg, adjust * wb, lo * ((wr + wg) + wb));
                wmax = fmaxf(wmax, fmaxf(fmaxf(fabsf(wv[i].x), fabsf(wv[i].y)), fmaxf(fabsf(wv[i].z), fabsf(wv[i].w))));
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) s_red[warp] = wmax;
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    pdl_trigger();                                                    // (TMEM is allocated)
    const uint32_t tmem = tmem_slot;
    wmax = s_red[0];
#pragma unroll
    for (int i = 1; i < 8; i++) wmax = fmaxf(wmax, s_red[i]);
    // ---- texel gather (the sampler of stem_kernel, unchanged), stored as FP16 quads; columns >= IW are padding for the
    // zero-weight tap slots and must hold finite values.  All of a thread's texel loads are issued first; the weight split
    // below runs while they are in flight ----
    constexpr int TPT = (IH * P + NT - 1) / NT;                       // texels per thread
    unsigned rgba[TPT];
    unsigned inside_mask = 0;
    if (separable) {
        const uint8_t *fbase = f.base + (long long)v.frame * f.frame_stride;
#pragma unroll
        for (int u = 0; u < TPT; u++) {
            const int e = tid + u * NT;
            rgba[u] = 0u;
            if (e < IH * P) {
                const int ty = e / P, tx = e - ty * P;
                const int sc = tx < IW ? s_col[tx] : -2, sr = s_row[ty];
                if (sc != -2 && sr != -2) {
                    inside_mask |= 1u << u;
                    if (sc >= 0 && sr >= 0)
                        rgba[u] = __ldg(reinterpret_cast<const unsigned *>(fbase + (long long)sr * f.row_stride + (long long)sc * 4));
                }
            }
        }
    } else {
#pragma unroll
        for (int u0 = 0; u0 < TPT; u0 += 4) {
            const unsigned *addr[4];
#pragma unroll
            for (int u = u0; u < u0 + 4 && u < TPT; u++) {
                const int e = tid + u * NT;
                const int ty = e / P, tx = e - ty * P;
                const int iy = iy_org + ty, ix = ix_org + tx;
                const bool in = e < IH * P && tx < IW && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                inside_mask |= in ? (1u << u) : 0u;
                addr[u - u0] = in ? sample_address(f, v, ix, iy, p.W, p.H) : nullptr;
            }
#pragma unroll
            for (int u = u0; u < u0 + 4 && u < TPT; u++) rgba[u] = addr[u - u0] ? __ldg(addr[u - u0]) : 0u;
        }
    }
    // power-of-two scale that puts the largest folded weight into [2^14, 2^15): FP16 hi + lo then carry 22 bits of every
    // weight that matters, and the products (x <= 255) stay far inside FP32
    int sexp = 0;
    if (wmax > 0.0f && wmax < 3.0e38f) sexp = 14 - ilogbf(wmax);
    sexp = max(-60, min(60, sexp));
    const float wscale = ldexpf(1.0f, sexp), inv_scale = ldexpf(1.0f, -sexp);
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const int e = tid + i * NT;
        if (e < ITEMS) {
            const int n = e % N, t = e / N, kx = t % KX, ky = t / KX;
            const float x0 = wv[i].x * wscale, x1 = wv[i].y * wscale, x2 = wv[i].z * wscale, x3 = wv[i].w * wscale;
            const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1), h2 = __float2half_rn(x2), h3 = __float2half_rn(x3);
            const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1)),
                         l2 = __float2half_rn(x2 - __half2float(h2)), l3 = __float2half_rn(x3 - __half2float(h3));
            const int off = ((ky * (KX / 2) + (kx >> 1)) * N + n) * 16 + (kx & 1) * 8;
            __half2 a = __halves2half2(h0, h1), c = __halves2half2(h2, h3);
            *reinterpret_cast<uint2 *>(s_B + G::PART_BYTES + off) = make_uint2(*reinterpret_cast<unsigned *>(&a), *reinterpret_cast<unsigned *>(&c));
            a = __halves2half2(l0, l1), c = __halves2half2(l2, l3);
            *reinterpret_cast<uint2 *>(s_B + off) = make_uint2(*reinterpret_cast<unsigned *>(&a), *reinterpret_cast<unsigned *>(&c));
        }
    }
#pragma unroll
    for (int u = 0; u < TPT; u++) {
        const int e = tid + u * NT;
        if (e < IH * P) s_tile[e] = f16 ? texel_f16_mapped(rgba[u], (inside_mask >> u) & 1u, lo, adjust) : texel_f16(rgba[u], (inside_mask >> u) & 1u);
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    // one issuing thread per accumulator (lane 0 of warps 0 .. jmax - 1): 2 * KS * KSTEPS MMAs each, descriptors = base + constants
    if (lane == 0 && warp < jmax) {
        tc::tc_fence_after();
        const int j = warp;
        const uint32_t idesc = make_idesc_f16(128, N);
        const uint64_t ad0 = tc::make_smem_desc(tc::smem_u32(s_tile) + (uint32_t)(j * 128), 16, 16 * P);
        const uint64_t bd0 = tc::make_smem_desc(tc::smem_u32(s_B), N * 16, 128);
        const uint32_t d = tmem + (uint32_t)(j * N);
#pragma unroll
        for (int part = 0; part < 2; part++) {                        // lo first, then hi
#pragma unroll
            for (int ky = 0; ky < KS; ky++) {
#pragma unroll
                for (int ks = 0; ks < G::KSTEPS; ks++) {
                    const uint64_t ad = ad0 + (uint64_t)((ky * P * 8 + ks * 32) >> 4);
                    const uint64_t bd = bd0 + (uint64_t)((part * G::PART_BYTES + (ky * (KX / 2) + 2 * ks) * N * 16) >> 4);
                    umma_f16(d, ad, bd, idesc, (part | ky | ks) != 0);
                }
            }
        }
        tc::umma_commit(&mbar);
    }
    if (lane == 0) tc::mbar_wait(&mbar, 0);                           // every MMA has read the tile and written TMEM
    __syncwarp();                                                     // (one poller per warp instead of 256 spinning threads)
    tc::tc_fence_after();
    // ---- epilogue: warp (q, g) owns TMEM lanes 32 q .. 32 q + 31 of two "units": accumulators 2 g and 2 g + 1 (N <= 32), or the
    // two 32-column halves of accumulator g (N = 64).  It stages the unit's 32 pixels in its own slice of shared memory (the
    // tile and the weights are dead) and writes them out as contiguous float4 runs: the 8 pixels of one output row of an MMA
    // tile are 8 * Ns consecutive floats of the NHWC tensor (Wo % 8 == 0) ----
    {
        constexpr int NLD = G::NLD;
        const int q = warp & 3, g = warp >> 2;
        float *stg = s_stage + warp * 32 * G::STAGE_STRIDE;
        const int act = p.epi.act1.kind;
        float *obase = p.out + (long long)img * p.out_img_stride;
        const int nqt = p.Ns >> 2;                                    // float4s per pixel of the output tensor
        for (int u = 0; u < 2; u++) {
            const int j = N <= 32 ? 2 * g + u : g, c0 = N <= 32 ? 0 : 32 * u;   // accumulator, first column of the unit
            if (j >= jmax || c0 >= p.Ns) break;
            float a[NLD];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * N + c0);
            if constexpr (NLD == 32) {
                tc::tmem_ld32(taddr, a);
            } else {
                tc::tmem_ld16(taddr, a);
            }
#pragma unroll
            for (int c = 0; c < NLD; c++) a[c] = fmaf(a[c], inv_scale, s_bias[c0 + c]);   // (explicit: this file is built with -fmad=false)
            if (act == ACT_RELU) {
#pragma unroll
                for (int c = 0; c < NLD; c++) a[c] = fmaxf(a[c], 0.0f);
            } else if (act == ACT_PRELU) {
#pragma unroll
                for (int c = 0; c < NLD; c++) a[c] = a[c] < 0.0f ? a[c] * s_sl[c0 + c] : a[c];
            } else if (act == ACT_CLIP) {
                const float clo = p.epi.act1.lo, chi = p.epi.act1.hi;
#pragma unroll
                for (int c = 0; c < NLD; c++) a[c] = fminf(fmaxf(a[c], clo), chi);
            }
#pragma unroll
            for (int c = 0; c < NLD; c += 4)
                *reinterpret_cast<float4 *>(stg + lane * G::STAGE_STRIDE + c) = make_float4(a[c], a[c + 1], a[c + 2], a[c + 3]);
            __syncwarp();
            auto write_out = [&](auto nq_tag) {
                constexpr int NQ = decltype(nq_tag)::value;           // float4s per pixel in this unit
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const int oy = oy0 + 4 * q + r;
                    if (oy >= p.Ho) break;
                    float4 *dst = reinterpret_cast<float4 *>(obase + ((long long)oy * p.Wo + ox0 + 8 * j) * p.Ns + c0);
#pragma unroll
                    for (int i0 = 0; i0 < 8 * NQ; i0 += 32) {
                        const int i = i0 + lane;
                        if ((8 * NQ) % 32 == 0 || i < 8 * NQ) {
                            const int px = i / NQ, c4 = i - px * NQ;
                            dst[N <= 32 ? i : px * nqt + c4] = *reinterpret_cast<const float4 *>(stg + (8 * r + px) * G::STAGE_STRIDE + 4 * c4);
                        }
                    }
                }
            };
            const int nq = min(NLD / 4, nqt - (c0 >> 2));
            if (nq == 4) write_out(std::integral_constant<int, 4>{});
            else if (nq == 6) write_out(std::integral_constant<int, 6>{});
            else if (nq == 8) write_out(std::integral_constant<int, 8>{});
            else if (nq == 2) write_out(std::integral_constant<int, 2>{});
            else if (nq == 3) write_out(std::integral_constant<int, 3>{});
            else if (nq == 5) write_out(std::integral_constant<int, 5>{});
            else if (nq == 7) write_out(std::integral_constant<int, 7>{});
            else write_out(std::integral_constant<int, 1>{});
            __syncwarp();
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, G::TCOLS);
}

template <int KS, int N, int MINB>
bool launch_stem_mma_cfg(const FramesDev &f, const ViewDev *views, float lo, float hi, const ConvDev &p, int f16, cudaStream_t s) {
    using G = StemMma<KS, N>;
    auto kern = stem_mma_kernel<KS, N, MINB>;
    static SmemOptIn opt_in;
    if (!opt_in.ensure(kern, G::SMEM_BYTES)) return false;
    const int tiles_x = (p.Wo + G::TW - 1) / G::TW, tiles_y = (p.Ho + G::TH - 1) / G::TH;
    const int images = p.M / (p.Ho * p.Wo);
    ZB_KNAME("stem_mma_kernel", KS, N, MINB);
    launch_pdl(16, kern, dim3((unsigned)(tiles_x * tiles_y * images)), dim3(256), G::SMEM_BYTES, s, f, views, lo, hi, p, tiles_x, tiles_y, f16);
    return true;
}

// ------------------------------------------------------------------------------------------------
// Detection decode for anchor i, field f:
//   0 cx, 1 cy, 2 w, 3 h, 4 angle, 5+2k kp_k.x, 6+2k kp_k.y   (network-input pixel units)
// ------------------------------------------------------------------------------------------------
struct AnchorXY { float x, y; };

__device__ __forceinline__ AnchorXY anchor_of(const DecodeParams &p, int i) {
    // Anchors::calculate (ssd.rs:96-119): layers in order, y-major, x, then boxes_per_cell copies.
    int w = p.l0_w, h = p.l0_h, b = p.l0_boxes;
    const int n0 = p.l0_w * p.l0_h * p.l0_boxes;
    if (i >= n0) {
        i -= n0;
        w = p.l1_w, h = p.l1_h, b = p.l1_boxes;
    }
    const int cell = i / b;
    const int cy = cell / w, cx = cell - cy * w;
    AnchorXY a;
    a.x = ((float)cx + 0.5f) / (float)w;
    a.y = ((float)cy + 0.5f) / (float)h;
    return a;
}

// Vec2::signed_angle_to (zaru-linalg vector.rs:568-573)
__device__ __forceinline__ float signed_angle_to(float ax, float ay, float bx, float by) {
    const float perp = ax * by - ay * bx;
    const float dot = (0.0f + ax * bx) + ay * by;
    return -(float)atan2((double)perp, (double)dot);   // glibc atan2f is correctly rounded in practice
}

__device__ float decode_field(const DecodeParams &p, const float *__restrict__ bp, int anchor, int field) {
    const AnchorXY a = anchor_of(p, anchor);
    const float isx = (float)p.net_w, isy = (float)p.net_h;
    const float cx = bp[0] + a.x * isx;
    const float cy = bp[1] + a.y * isy;
    if (field == 0) return cx;
    if (field == 1) return cy;
    if (field == 2) return bp[2];
    if (field == 3) return bp[3];
    if (field == 4) {
        // keypoint k = (b[4+2k], b[5+2k]) + center * input_size   (center already in pixels: SURVEY F4)
        if (p.angle_kind == 0) {
            const float lx = bp[4] + cx * isx, ly = bp[5] + cy * isy;
            const float rx = bp[6] + cx * isx, ry = bp[7] + cy * isy;
            return signed_angle_to(rx - lx, ry - ly, 1.0f, 0.0f);
        }
        const float wx = bp[4] + cx * isx, wy = bp[5] + cy * isy;      // Wrist = kp 0
        const float fx = bp[8] + cx * isx, fy = bp[9] + cy * isy;      // MiddleFingerMcp = kp 2
        return signed_angle_to(wx - fx, wy - fy, 0.0f, 1.0f);
    }
    const int k = field - 5;
    return (k & 1) ? (bp[4 + k] + cy * isy) : (bp[4 + k] + cx * isx);
}

// f32::total_cmp key (zaru-image/src/num.rs:5-28)
__device__ __forceinline__ int total_key(float v) {
    int b = __float_as_int(v);
    return b ^ (int)(((unsigned)(b >> 31)) >> 1);
}

// Rect::iou (rect.rs:193-214) on (cx, cy, w, h)
__device__ __forceinline__ float rect_area_span(float minx, float miny, float maxx, float maxy) {
    // Rect::bounding([min,max]) -> span_inner -> from_top_left(x_min, y_min, x_max - x_min, y_max - y_min); area = w*h
    return (maxx - minx) * (maxy - miny);
}
__device__ __forceinline__ float iou_ref(float4 a, float4 b) {
    const float ax = a.x - a.z * 0.5f, ay = a.y - a.w * 0.5f;
    const float bx = b.x - b.z * 0.5f, by = b.y - b.w * 0.5f;
    const float minx = fmaxf(ax, bx), miny = fmaxf(ay, by);
    const float maxx = fminf(ax + a.z, bx + b.z), maxy = fminf(ay + a.w, by + b.w);
    float inter = 0.0f;
    if (!(minx > maxx || miny > maxy)) inter = rect_area_span(minx, miny, maxx, maxy);
    const float uni = a.z * a.w + b.z * b.w - inter;
    return inter / uni;
}

// Order-preserving block compaction helper: returns the exclusive prefix of `flag` over the block and the total.
__device__ __forceinline__ int block_excl_scan(int flag, int *warp_sums, int &total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    const int in_warp = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_sums[wid] = __popc(bal);
    __syncthreads();
    int base = 0, tot = 0;
    for (int w = 0; w < nw; w++) {
        const int s = warp_sums[w];
        if (w < wid) base += s;
        tot += s;
    }
    __syncthreads();
    total = tot;
    return base + in_warp;
}

// One CTA per image.  Dynamic smem: A * (int idx + float conf + float4 box + int order + int rem + int tmp)
__global__ void __launch_bounds__(256) decode_nms_kernel(const float *__restrict__ boxes,
                                                         const float *__restrict__ scores,
                                                         const float *__restrict__ fit, const DecodeParams p,
                                                         DetDev *__restrict__ out, int *__restrict__ out_counts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int A = p.num_anchors;
    float4 *c_box = reinterpret_cast<float4 *>(smem_raw);
    float *c_conf = reinterpret_cast<float *>(c_box + A);
    int *c_idx = reinterpret_cast<int *>(c_conf + A);
    int *rem = c_idx + A;      // remaining candidates, ascending (confidence, index)
    int *tmp = rem + A;        // scratch for compaction
    int *members = tmp + A;    // cluster members (ascending), seed excluded
    __shared__ int warp_sums[8];

    const int img = blockIdx.x;
    const float *bx = boxes + (long long)img * A * p.num_params;
    const float *sc = scores + (long long)img * A;
    const int tid = threadIdx.x, T = blockDim.x;
    // Detector::timers() split (detection.rs:155-157, :231-243): when asked for, the CTA's extract / nms phase durations
    // are summed (ns) so the host can apportion the kernel's event time between t_extract and t_nms
    unsigned long long t_start = 0, t_mid = 0;
    if (p.phase_ns && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));

    // 1. sigmoid + threshold, compact in anchor order  (face/detection.rs:109-113: `if conf < thresh {continue}`)
    int n = 0;
    for (int base = 0; base < A; base += T) {
        const int i = base + tid;
        float conf = 0.f;
        int keep = 0;
        if (i < A) {
            conf = sigmoid_ref(sc[i]);
            keep = !(conf < p.thresh);
        }
        int tot;
        const int pos = block_excl_scan(keep, warp_sums, tot);
        if (keep) {
            c_idx[n + pos] = i;
            c_conf[n + pos] = conf;
            const float *bp = bx + (long long)i * p.num_params;
            c_box[n + pos] = make_float4(decode_field(p, bp, i, 0), decode_field(p, bp, i, 1), bp[2], bp[3]);
        }
        n += tot;
    }
    __syncthreads();
    if (p.phase_ns && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_mid));

    // 2. stable ascending rank sort by TotalF32(confidence)  (nms.rs:66; tie order fixed as stable)
    for (int c = tid; c < n; c += T) {
        const int kc = total_key(c_conf[c]);
        int rank = 0;
        for (int d = 0; d < n; d++) {
            const int kd = total_key(c_conf[d]);
            rank += (kd < kc) || (kd == kc && d < c);
        }
        rem[rank] = c;
    }
    __syncthreads();

    // 3. greedy seed loop
    const float scale = fit[img * 4 + 0], tlx = fit[img * 4 + 1], tly = fit[img * 4 + 2];
    const int nfields = 5 + 2 * p.num_kp;
    int n_rem = n, n_out = 0;
    while (n_rem > 0) {
        const int seed = rem[n_rem - 1];
        const float4 sbox = c_box[seed];
        n_rem -= 1;
        __syncthreads();
        // classify the remaining candidates against the seed, preserving order
        int n_keep = 0, n_mem = 0;
        for (int base = 0; base < n_rem; base += T) {
            const int j = base + tid;
            int c = -1, is_mem = 0, is_keep = 0;
            if (j < n_rem) {
                c = rem[j];
                const float iou = iou_ref(sbox, c_box[c]);
                if (p.nms_mode == 1) {
                    is_mem = iou >= p.iou_thresh;      // nms.rs:86
                    is_keep = !is_mem;
                } else {
                    is_keep = iou < p.iou_thresh;      // nms.rs:73 (NaN is dropped)
                }
            }
            int tk, tm;
            const int pk = block_excl_scan(is_keep, warp_sums, tk);
            const int pm = block_excl_scan(is_mem, warp_sums, tm);
            if (is_keep) tmp[n_keep + pk] = c;
            if (is_mem) members[n_mem + pm] = c;
            n_keep += tk;
            n_mem += tm;
        }
        __syncthreads();
        for (int j = tid; j < n_keep; j += T) rem[j] = tmp[j];
        n_rem = n_keep;

        // output
        if (n_out < p.cap) {
            DetDev *o = out + (long long)img * p.cap + n_out;
            if (tid < nfields) {
                float val;
                const float *sbp = bx + (long long)c_idx[seed] * p.num_params;
                if (p.nms_mode == 1) {
                    // confidence-weighted average over [seed, members ascending]  (nms.rs:93-131)
                    float acc = 0.0f, divisor = 0.0f;
                    {
                        const float factor = c_conf[seed];
                        divisor = divisor + factor;
                        acc = acc + decode_field(p, sbp, c_idx[seed], tid) * factor;
                    }
                    for (int j = 0; j < n_mem; j++) {
                        const int c = members[j];
                        const float factor = c_conf[c];
                        divisor = divisor + factor;
                        acc = acc + decode_field(p, bx + (long long)c_idx[c] * p.num_params, c_idx[c], tid) * factor;
                    }
                    val = acc / divisor;
                } else {
                    val = decode_field(p, sbp, c_idx[seed], tid);
                }
                // remap to the caller's coordinates (detection.rs:245-267)
                if (tid == 0) o->cx = val * scale + tlx;
                else if (tid == 1) o->cy = val * scale + tly;
                else if (tid == 2) o->w = val * scale;
                else if (tid == 3) o->h = val * scale;
                else if (tid == 4) o->angle = val;
                else {
                    const int k = tid - 5;
                    o->kp[k] = val * scale + ((k & 1) ? tly : tlx);
                }
            } else if (tid == nfields) {
                o->confidence = c_conf[seed];
                o->num_kp = p.num_kp;
                o->anchor = c_idx[seed];
            } else if (tid > nfields && tid <= nfields + (14 - 2 * p.num_kp)) {
                o->kp[2 * p.num_kp + (tid - nfields - 1)] = 0.0f;
            }
        }
        n_out += 1;
        __syncthreads();
    }
    if (tid == 0) out_counts[img] = n_out;
    if (p.phase_ns && tid == 0) {
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        atomicAdd(p.phase_ns, t_mid - t_start);
        atomicAdd(p.phase_ns + 1, t_end - t_mid);
    }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cos_sin_ref(float rad, float &c, float &s);
__global__ void __launch_bounds__(128) face_roi_kernel(const FramesDev f, const DetDev *__restrict__ dets,
                                                       const int *__restrict__ counts, int cap, int first_frame,
                                                       int n, int net_w, int net_h, ViewDev *__restrict__ out_views,
                                                       float *__restrict__ out_fit, ViewHost *__restrict__ out_rects,
                                                       float grow, int use_angle) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cnt = min(counts[i], cap);
    ViewDev v;
    v.frame = first_frame + i;
    v.flip_x = 0;
    v.valid = cnt > 0;
    v.cx = v.cy = 0.f, v.w = v.h = 1.f, v.cosr = 1.f, v.sinr = 0.f;
    float fit0 = 1.f, fit1 = 0.f, fit2 = 0.f;
    ViewHost vr;
    vr.frame = first_frame + i;
    vr.cx = vr.cy = vr.w = vr.h = 0.f, vr.radians = 0.f;
    if (cnt > 0) {
        // detections.iter().max_by_key(|det| TotalF32(det.confidence()))  -> last maximum (facemesh.rs:49-52)
        const DetDev *d = dets + (long long)i * cap;
        int best = 0;
        int bk = __float_as_int(d[0].confidence);
        bk ^= (int)(((unsigned)(bk >> 31)) >> 1);
        for (int j = 1; j < cnt; j++) {
            int k = __float_as_int(d[j].confidence);
            k ^= (int)(((unsigned)(k >> 31)) >> 1);
            if (k >= bk) bk = k, best = j;
        }
        const float aspect = aspect_as_f32((unsigned)net_w, (unsigned)net_h);
        // tracker.set_roi(detection.bounding_rect()); view_rect = roi.map(grow_to_fit_aspect)  (landmark.rs:465)
        RectF roi;
        roi.cx = d[best].cx, roi.cy = d[best].cy, roi.w = d[best].w, roi.h = d[best].h;
        // hand: RotatedRect::new(det.bounding_rect().grow_rel(1.5), det.angle())   (hand/tracking.rs:136, :159)
        if (grow != 0.0f) roi = grow_rel(roi, grow);
        const float rad = use_angle ? d[best].angle : 0.0f;
        float c, s;
        cos_sin_ref(rad, c, s);                  // rotation 0 -> cos 1, sin 0 exactly
        const RectF view_rect = grow_to_fit_aspect(roi, aspect);
        vr.cx = view_rect.cx, vr.cy = view_rect.cy, vr.w = view_rect.w, vr.h = view_rect.h, vr.radians = rad;
        // full_image.view(view_rect)   (landmark.rs:466)
        const RRectF full = full_view(f.width, f.height);
        const RRectF v1 = view_compose(full, view_rect, rad, c, s);
        // Estimator::estimate_impl: rect = view.rect().grow_to_fit_aspect(..); view = image.view(rect)  (:320-323)
        const RectF r1 = rect_from_top_left(0.0f, 0.0f, v1.r.w, v1.r.h);
        const RectF r2 = grow_to_fit_aspect(r1, aspect);
        const RRectF v2 = view_compose(v1, r2, 0.0f, c, s);
        v.cx = v2.r.cx, v.cy = v2.r.cy, v.w = v2.r.w, v.h = v2.r.h, v.cosr = c, v.sinr = s;
        fit0 = r2.w / (float)net_w;   // scale = rect.width() / input_res.width()   (:336)
        fit1 = rect_x(r2);
        fit2 = rect_y(r2);
    }
    out_views[i] = v;
    out_fit[i * 4 + 0] = fit0, out_fit[i * 4 + 1] = fit1, out_fit[i * 4 + 2] = fit2, out_fit[i * 4 + 3] = 0.f;
    if (out_rects) out_rects[i] = vr;
}

// ------------------------------------------------------------------------------------------------
// zaru::filter, one scalar: st = (has, x|last, dx|v).  f32 operation order of the reference; this TU has FMA
// contraction off, so results are bit-identical to the scalar Rust code.
__device__ __forceinline__ float smoothing_factor(float t_e, float cutoff) {      // one_euro.rs:90-93
    const float r = 2.0f * 3.14159265358979323846f * cutoff * t_e;
    return r / (r + 1.0f);
}
__device__ __forceinline__ float filter_scalar(const FilterDev &f, float *st, float x) {
    if (f.kind == FILTER_EMA) {                                                    // ema.rs:29-42
        if (st[0] != 0.0f) {
            const float avg = f.p0 * x + (1.0f - f.p0) * st[1];
            st[1] = avg;
            return avg;
        }
        st[0] = 1.0f, st[1] = x;
        return x;
    }
    if (f.kind == FILTER_ONE_EURO) {                                               // one_euro.rs:63-87
        if (st[0] == 0.0f) {
            st[0] = 1.0f, st[1] = x, st[2] = 0.0f;
            return x;
        }
        const float a_d = smoothing_factor(f.elapsed, f.p2);
        const float dx = (x - st[1]) / f.elapsed;
        const float dx_hat = a_d * dx + (1.0f - a_d) * st[2];
        const float cutoff = f.p0 + f.p1 * fabsf(dx_hat);
        const float a = smoothing_factor(f.elapsed, cutoff);
        const float x_hat = a * x + (1.0f - a) * st[1];
        st[1] = x_hat, st[2] = dx_hat;
        return x_hat;
    }
    if (f.kind == FILTER_ALPHA_BETA) {                                             // alpha_beta.rs:31-49
        if (st[0] == 0.0f) {
            st[0] = 1.0f, st[1] = x;
            return x;
        }
        const float prediction = st[1] + st[2] * f.elapsed;
        const float residual = x - prediction;
        st[1] = prediction + f.p0 * residual;
        st[2] = st[2] + f.p1 * residual / f.elapsed;
        return st[1];
    }
    return x;
}

__global__ void __launch_bounds__(256) filter_apply_kernel(const FilterDev f, float *__restrict__ values, long long count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) values[i] = filter_scalar(f, f.state + 3 * i, values[i]);
}

// Frames whose RoI view is valid (a face / palm was detected), in frame order: sel[j] = frame, out_views[j] = its view,
// *count = how many.  The landmark network then runs on `count` images instead of all n - the reference only calls its
// estimator when the detector found something (examples/facemesh.rs:49-55).  Frames without a detection get their
// result scalars here (flag / presence -1, second scalar 0); their landmark rows are zero-filled by the caller.
__global__ void __launch_bounds__(1024) compact_views_kernel(const ViewDev *__restrict__ views, int n, ViewDev *__restrict__ out_views,
                                                             int *__restrict__ sel, int *__restrict__ count,
                                                             float *__restrict__ scalars) {
    __shared__ int warp_sums[32];
    __shared__ int base;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const bool v = i < n && views[i].valid;
        const unsigned m = __ballot_sync(0xffffffffu, v);
        if (lane == 0) warp_sums[w] = __popc(m);
        __syncthreads();
        int off = base;
        for (int k = 0; k < w; k++) off += warp_sums[k];
        if (v) {
            const int j = off + __popc(m & ((1u << lane) - 1u));
            sel[j] = i;
            out_views[j] = views[i];
        } else if (i < n && scalars) {
            scalars[2 * i + 0] = -1.0f;
            scalars[2 * i + 1] = 0.0f;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int k = 0; k < 32; k++) t += warp_sums[k];
            base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}

__global__ void __launch_bounds__(128) landmarks_kernel(const float *__restrict__ out0, int s0,
                                                        const float *__restrict__ out1, int s1,
                                                        const float *__restrict__ out2, int s2,
                                                        const float *__restrict__ fit,
                                                        const ViewDev *__restrict__ views,
                                                        const ViewHost *__restrict__ view_rects, int n,
                                                        const LandmarkParams p, float *__restrict__ landmarks,
                                                        float *__restrict__ scalars, const FilterDev flt,
                                                        const int *__restrict__ sel) {
    // `sel` (compacted pipeline): network output row blockIdx.y belongs to frame sel[blockIdx.y]; the raw tensors are
    // indexed by the row, everything per frame (fit, views, view rects, filter state, results) by the frame
    // (the batch index lives on grid.x: grid.y is limited to 65535)
    const int bpi = (p.num_landmarks + (int)blockDim.x - 1) / (int)blockDim.x;
    const int brow = (int)(blockIdx.x / bpi);
    const int img = sel ? sel[brow] : brow;
    const int l = (int)(blockIdx.x - (unsigned)brow * bpi) * blockDim.x + threadIdx.x;
    const int valid = views ? views[img].valid : 1;
    if (sel) {
        const long long row = brow;
        out0 += (row - img) * s0;
        if (out1) out1 += (row - img) * s1;
        if (out2) out2 += (row - img) * s2;
    }
    if (l == 0 && scalars) {
        float a = 0.f, b = 0.f;
        if (p.kind == 0) a = valid ? sigmoid_ref(out1[(long long)img * s1]) : -1.0f;  // mediapipe.rs:60
        if (p.kind == 3) {   // FaceMeshV2 (mediapipe.rs:96-99): sigmoid(face flag); tongueOut is already a probability
            a = valid ? sigmoid_ref(out1[(long long)img * s1]) : -1.0f;
            b = out2[(long long)img * s2];
        }
        if (p.kind == 2) {                                                             // hand/landmark.rs:310-311
            a = valid ? out1[(long long)img * s1] : -1.0f;
            b = valid ? out2[(long long)img * s2] : 0.0f;
        }
        scalars[img * 2 + 0] = a, scalars[img * 2 + 1] = b;
    }
    if (l >= p.num_landmarks) return;
    const float *src;
    if (p.kind == 1) {
        // eye.rs:51-63: iris (out1, 5 pts) -> positions[..5]; contour (out0, 71 pts) -> positions[5..]
        src = (l < 5) ? out1 + (long long)img * s1 + 3 * l : out0 + (long long)img * s0 + 3 * (l - 5);
    } else {
        src = out0 + (long long)img * s0 + 3 * l;
    }
    float x = src[0], y = src[1], z = src[2];
    if (flt.state && flt.kind != FILTER_NONE && valid) {
        // self.filter.filter(self.estimate.landmarks_mut()) in network coordinates (landmark.rs:330-333)
        float *st = flt.state + ((long long)img * p.num_landmarks + l) * 9;
        x = filter_scalar(flt, st, x);
        y = filter_scalar(flt, st + 3, y);
        z = filter_scalar(flt, st + 6, z);
        // the estimate now holds the filtered positions: later readers of the raw tensor (the tracker's
        // angle_radians on the eye corners) must see them too
        float *w = const_cast<float *>(src);
        w[0] = x, w[1] = y, w[2] = z;
    }
    if (views && views[img].flip_x) {
        // EyeLandmarks::flip_horizontal_in_place (eye.rs:121-125) in network-input coordinates
        const float half = (float)p.net_w / 2.0f;
        x = -(x - half) + half;
    }
    const float scale = fit[img * 4 + 0], tlx = fit[img * 4 + 1], tly = fit[img * 4 + 2];
    x = x * scale, y = y * scale, z = z * scale;   // landmark.rs:339 (z is scaled too)
    x = x + tlx;
    y = y + tly;
    if (p.track_transform && view_rects) {
        // LandmarkTracker::track: view_rect.transform_out([p.x, p.y])  (landmark.rs:482-486)
        const ViewHost vr = view_rects[img];
        RRectF rr;
        rr.r.cx = vr.cx, rr.r.cy = vr.cy, rr.r.w = vr.w, rr.r.h = vr.h;
        rr.rad = vr.radians;
        rr.c = (vr.radians == 0.0f) ? 1.0f : (float)cos((double)vr.radians);
        rr.s = (vr.radians == 0.0f) ? 0.0f : (float)sin((double)vr.radians);
        float ox, oy;
        transform_out(rr, x, y, ox, oy);
        x = ox, y = oy;
    }
    float *o = landmarks + ((long long)img * p.num_landmarks + l) * 3;
    if (!valid) x = y = z = 0.f;
    o[0] = x, o[1] = y, o[2] = z;
}

// ------------------------------------------------------------------------------------------------
// LandmarkTracker::track_impl (landmark.rs:463-501), device resident.
// cos/sin of a device-computed angle: evaluated in f64 and rounded once (glibc cosf/sinf are correctly rounded in
// practice, CUDA's are not), the same convention as exp/atan2 above.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cos_sin_ref(float rad, float &c, float &s) {
    if (rad == 0.0f) {
        c = 1.0f, s = 0.0f;
    } else {
        c = (float)cos((double)rad), s = (float)sin((double)rad);
    }
}

// roi -> (optional) grow_to_fit_aspect -> full_image.view(rect) -> Estimator view fit.  pre_fit = 1 is the tracker's
// `view_rect = roi.map(grow_to_fit_aspect)` (landmark.rs:465); pre_fit = 0 is `estimator.estimate(&image.view(roi))`.
__device__ __forceinline__ void roi_to_view(const FramesDev &f, int frame, RectF roi, float rad, int pre_fit, int net_w, int net_h,
                                            ViewDev &v, float &fit0, float &fit1, float &fit2, ViewHost &vr) {
    const float aspect = aspect_as_f32((unsigned)net_w, (unsigned)net_h);
    // let view_rect = roi.map(|rect| rect.grow_to_fit_aspect(self.aspect_ratio));   (landmark.rs:465)
    const RectF view_rect = pre_fit ? grow_to_fit_aspect(roi, aspect) : roi;
    vr.frame = frame;
    vr.cx = view_rect.cx, vr.cy = view_rect.cy, vr.w = view_rect.w, vr.h = view_rect.h, vr.radians = rad;
    float c, s;
    cos_sin_ref(rad, c, s);
    // full_image.view(view_rect)   (landmark.rs:466; parent rotation 0, so the summed angle is roi's)
    const RRectF full = full_view(f.width, f.height);
    const RRectF v1 = view_compose(full, view_rect, rad, c, s);
    // Estimator::estimate_impl: rect = view.rect().grow_to_fit_aspect(..); view = image.view(rect)  (:320-323)
    const RectF r1 = rect_from_top_left(0.0f, 0.0f, v1.r.w, v1.r.h);
    const RectF r2 = grow_to_fit_aspect(r1, aspect);
    const RRectF v2 = view_compose(v1, r2, 0.0f, c, s);
    v.frame = frame;
    v.valid = 1;
    v.cx = v2.r.cx, v.cy = v2.r.cy, v.w = v2.r.w, v.h = v2.r.h, v.cosr = c, v.sinr = s;
    fit0 = r2.w / (float)net_w;
    fit1 = rect_x(r2);
    fit2 = rect_y(r2);
}

__global__ void __launch_bounds__(128) tracker_prepare_kernel(const FramesDev f, const TrackState *__restrict__ state,
                                                              int first_frame, int n, int net_w, int net_h,
                                                              ViewDev *__restrict__ out_views, float *__restrict__ out_fit,
                                                              ViewHost *__restrict__ out_rects) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const TrackState st = state[i];
    ViewDev v;
    v.frame = first_frame + i;
    v.flip_x = 0;
    v.valid = 0;
    v.cx = v.cy = 0.f, v.w = v.h = 1.f, v.cosr = 1.f, v.sinr = 0.f;
    float fit0 = 1.f, fit1 = 0.f, fit2 = 0.f;
    ViewHost vr;
    vr.frame = first_frame + i;
    vr.cx = vr.cy = vr.w = vr.h = 0.f, vr.radians = 0.f;
    if (st.has) {
        RectF roi;
        roi.cx = st.cx, roi.cy = st.cy, roi.w = st.w, roi.h = st.h;
        roi_to_view(f, first_frame + i, roi, st.rad, 1, net_w, net_h, v, fit0, fit1, fit2, vr);
    }
    out_views[i] = v;
    out_fit[i * 4 + 0] = fit0, out_fit[i * 4 + 1] = fit1, out_fit[i * 4 + 2] = fit2, out_fit[i * 4 + 3] = 0.f;
    out_rects[i] = vr;
}

// Caller-supplied RoIs (rois[i].frame names the frame): one tracker-style step per RoI, as tracker_prepare_kernel does
// for the RoIs it holds itself.
__global__ void __launch_bounds__(128) rois_prepare_kernel(const FramesDev f, const ViewHost *__restrict__ rois, int n, int net_w,
                                                           int net_h, ViewDev *__restrict__ out_views, float *__restrict__ out_fit,
                                                           ViewHost *__restrict__ out_rects) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ViewHost r = rois[i];
    ViewDev v;
    v.flip_x = 0;
    float fit0, fit1, fit2;
    ViewHost vr;
    RectF roi;
    roi.cx = r.cx, roi.cy = r.cy, roi.w = r.w, roi.h = r.h;
    roi_to_view(f, r.frame, roi, r.radians, 1, net_w, net_h, v, fit0, fit1, fit2, vr);
    out_views[i] = v;
    out_fit[i * 4 + 0] = fit0, out_fit[i * 4 + 1] = fit1, out_fit[i * 4 + 2] = fit2, out_fit[i * 4 + 3] = 0.f;
    out_rects[i] = vr;
}

// LandmarkResultV1::left_eye() / right_eye() (mediapipe.rs:146-192) for face i -> eye views 2i (left) and 2i + 1 (right,
// mirrored: eye.rs:24-28) of BASELINE config 2:
//   rotation = (right_eye_outer - left_eye_outer).signed_angle_to(X)
//   eye      = RotatedRect::bounding(rotation, [bottom, corner, corner, top])       (rect.rs:287-325)
//   view     = image.view(eye.map(|r| r.grow_rel(margin)))  ->  Estimator::estimate(&view)   (landmark.rs:314-348)
// `landmarks` are the face-mesh positions in FRAME coordinates ([n][L][3]); face_views[i] gives the frame and validity.
__global__ void __launch_bounds__(128) eye_roi_kernel(const FramesDev f, const float *__restrict__ landmarks,
                                                      const ViewDev *__restrict__ face_views, int n, int L, int net_w, int net_h,
                                                      float margin, ViewDev *__restrict__ out_views, float *__restrict__ out_fit,
                                                      ViewHost *__restrict__ out_rects) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 2 * n) return;
    const int i = e >> 1, right = e & 1;
    const ViewDev fv = face_views[i];
    ViewDev v;
    v.frame = fv.frame;
    v.flip_x = right;
    v.valid = 0;
    v.cx = v.cy = 0.f, v.w = v.h = 1.f, v.cosr = 1.f, v.sinr = 0.f;
    float fit0 = 1.f, fit1 = 0.f, fit2 = 0.f;
    ViewHost vr;
    vr.frame = fv.frame;
    vr.cx = vr.cy = vr.w = vr.h = 0.f, vr.radians = 0.f;
    if (fv.valid) {
        const float *lm = landmarks + (long long)i * L * 3;
        // LandmarkIdx (mediapipe.rs:530-545)
        const float lx = lm[3 * 33], ly = lm[3 * 33 + 1], rx = lm[3 * 263], ry = lm[3 * 263 + 1];
        const float rot = signed_angle_to(rx - lx, ry - ly, 1.0f, 0.0f);
        const int idx_l[4] = {145, 33, 133, 159}, idx_r[4] = {374, 362, 263, 386};
        float c, s;
        cos_sin_ref(-rot, c, s);                       // cw = rotation_clockwise(rot) = ccw(-rot)
        float mnx = 3.402823466e+38f, mny = 3.402823466e+38f, mxx = -3.402823466e+38f, mxy = -3.402823466e+38f;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int l = right ? idx_r[k] : idx_l[k];
            float px, py;
            rot_ccw_apply(c, s, lm[3 * l], lm[3 * l + 1], px, py);
            mnx = fminf(mnx, px), mny = fminf(mny, py), mxx = fmaxf(mxx, px), mxy = fmaxf(mxy, py);
        }
        const float ccx = (mnx + mxx) * 0.5f, ccy = (mny + mxy) * 0.5f;
        float c2, s2, ox, oy;
        cos_sin_ref(rot, c2, s2);
        rot_ccw_apply(c2, s2, ccx, ccy, ox, oy);
        RectF roi;
        roi.cx = ox, roi.cy = oy, roi.w = mxx - mnx, roi.h = mxy - mny;
        if (margin != 0.0f) roi = grow_rel(roi, margin);
        roi_to_view(f, fv.frame, roi, rot, 0, net_w, net_h, v, fit0, fit1, fit2, vr);
        v.flip_x = right;
    }
    out_views[e] = v;
    out_fit[e * 4 + 0] = fit0, out_fit[e * 4 + 1] = fit1, out_fit[e * 4 + 2] = fit2, out_fit[e * 4 + 3] = 0.f;
    out_rects[e] = vr;
}

// One CTA per stream.
__global__ void __launch_bounds__(128) tracker_update_kernel(TrackState *__restrict__ state, const float *__restrict__ out0,
                                                             int s0, const float *__restrict__ fit,
                                                             const float *__restrict__ landmarks,
                                                             const float *__restrict__ scalars, int num_landmarks,
                                                             float loss_thresh, float roi_padding, int le_idx, int re_idx,
                                                             float axis_x, float axis_y,
                                                             ViewHost *__restrict__ out_updated,
                                                             unsigned char *__restrict__ out_tracked) {
    const int i = blockIdx.x, tid = threadIdx.x;
    __shared__ float s_red[4][4];
    const TrackState st = state[i];
    ViewHost up;
    up.frame = i;
    up.cx = up.cy = up.w = up.h = up.radians = 0.f;
    // `let roi = self.roi?;` and `if estimate.confidence() < self.loss_thresh { self.roi = None; return None; }`
    const bool lost = !st.has || scalars[i * 2] < loss_thresh;
    if (lost) {
        if (tid == 0) {
            state[i].has = 0;
            out_updated[i] = up;
            out_tracked[i] = 0;
        }
        return;
    }
    // estimate.angle_radians() in VIEW coordinates (positions are mapped to the image afterwards):
    //   face mesh: (right_eye_outer - left_eye_outer).signed_angle_to(X)        (mediapipe.rs:146-160)
    //   hand:      (wrist - middle_finger_mcp).signed_angle_to(Y)              (hand/landmark.rs:68-78)
    const float scale = fit[i * 4 + 0], tlx = fit[i * 4 + 1], tly = fit[i * 4 + 2];
    const float *o = out0 + (long long)i * s0;
    const float lx = o[3 * le_idx] * scale + tlx, ly = o[3 * le_idx + 1] * scale + tly;
    const float rx = o[3 * re_idx] * scale + tlx, ry = o[3 * re_idx + 1] * scale + tly;
    const float angle = st.rad + signed_angle_to(rx - lx, ry - ly, axis_x, axis_y);
    // RotatedRect::bounding(angle, points)  (rect.rs:287-325): cw = rotation_clockwise(angle) = ccw(-angle)
    float c, s;
    cos_sin_ref(-angle, c, s);
    float mnx = 3.402823466e+38f, mny = 3.402823466e+38f, mxx = -3.402823466e+38f, mxy = -3.402823466e+38f;
    const float *lm = landmarks + (long long)i * num_landmarks * 3;
    for (int l = tid; l < num_landmarks; l += blockDim.x) {
        float px, py;
        rot_ccw_apply(c, s, lm[3 * l], lm[3 * l + 1], px, py);
        mnx = fminf(mnx, px), mny = fminf(mny, py), mxx = fmaxf(mxx, px), mxy = fmaxf(mxy, py);
    }
    for (int d = 16; d > 0; d >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, d));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, d));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, d));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, d));
    }
    if ((tid & 31) == 0) s_red[tid >> 5][0] = mnx, s_red[tid >> 5][1] = mny, s_red[tid >> 5][2] = mxx, s_red[tid >> 5][3] = mxy;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) {
            mnx = fminf(mnx, s_red[w][0]), mny = fminf(mny, s_red[w][1]);
            mxx = fmaxf(mxx, s_red[w][2]), mxy = fmaxf(mxy, s_red[w][3]);
        }
        const float ccx = (mnx + mxx) * 0.5f, ccy = (mny + mxy) * 0.5f;     // centre in the rotated frame
        float c2, s2, ox, oy;
        cos_sin_ref(angle, c2, s2);
        rot_ccw_apply(c2, s2, ccx, ccy, ox, oy);                               // center.rotate_counterclockwise(angle)
        up.cx = ox, up.cy = oy, up.w = mxx - mnx, up.h = mxy - mny, up.radians = angle;
        out_updated[i] = up;
        out_tracked[i] = 1;
        RectF r;
        r.cx = ox, r.cy = oy, r.w = up.w, r.h = up.h;
        r = grow_rel(r, roi_padding);                                          // self.roi = updated.grow_rel(padding)
        TrackState ns;
        ns.cx = r.cx, ns.cy = r.cy, ns.w = r.w, ns.h = r.h, ns.rad = angle, ns.has = 1;
        state[i] = ns;
    }
}

__global__ void tracker_set_roi_kernel(TrackState *__restrict__ state, const int *__restrict__ ids,
                                       const ViewHost *__restrict__ rois, int k) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    TrackState st;
    st.cx = st.cy = st.w = st.h = st.rad = 0.f, st.has = 0;
    if (rois) st.cx = rois[j].cx, st.cy = rois[j].cy, st.w = rois[j].w, st.h = rois[j].h, st.rad = rois[j].radians, st.has = 1;
    state[ids[j]] = st;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// zaru_image::blend(&mut dest_view, &src_view)  (zaru-image/src/blend.rs:13-32, :44-93; view.rs:81-119; blend.wgsl:27-38;
// gpu.rs:191-205): the destination quad and the source UV rectangle are axis-aligned between the transformed top-left and
// bottom-right corners of the two views; linear filtering (ClampToEdge) in linear light on sRGB texels, UV outside [0,1]
// -> Color::NONE, replace (no real blending).  One thread per destination pixel of the quad's bounding box.
// Rules the wgpu backend leaves to the GPU are fixed as the APIs specify them (oracle/blend.py): pixel centres at +0.5,
// top-left fill rule, exact f32 weights, sRGB transfer in f64.
// ------------------------------------------------------------------------------------------------
struct BlendJob {
    int dframe, sframe;
    float dx0, dy0, dx1, dy1;      // destination: transformed (0,0) and (w,h) of the destination view
    float sx0, sy0, sx1, sy1;      // source: the same for the source view
    int bx, by, bw, bh;            // destination pixel box to visit
};

__constant__ float c_srgb_lut[256];

__device__ __forceinline__ unsigned char srgb_encode_ref(float lin) {
    double x = fmin(fmax((double)lin, 0.0), 1.0);
    const double sv = x <= 0.0031308 ? x * 12.92 : 1.055 * pow(x, 1.0 / 2.4) - 0.055;
    return (unsigned char)floor(sv * 255.0 + 0.5);
}

__global__ void __launch_bounds__(256) blend_kernel(const FramesDev dst, const FramesDev src, const BlendJob *__restrict__ jobs) {
    const BlendJob j = jobs[blockIdx.z];
    const int x = j.bx + blockIdx.x * 32 + (threadIdx.x & 31), y = j.by + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= j.bx + j.bw || y >= j.by + j.bh || x < 0 || y < 0 || x >= dst.width || y >= dst.height) return;
    const float cx = (float)x + 0.5f, cy = (float)y + 0.5f;
    const float xlo = fminf(j.dx0, j.dx1), xhi = fmaxf(j.dx0, j.dx1), ylo = fminf(j.dy0, j.dy1), yhi = fmaxf(j.dy0, j.dy1);
    if (!(xlo <= cx && cx < xhi && ylo <= cy && cy < yhi)) return;
    const float tx = (cx - j.dx0) / (j.dx1 - j.dx0), ty = (cy - j.dy0) / (j.dy1 - j.dy0);
    const float px = j.sx0 + tx * (j.sx1 - j.sx0), py = j.sy0 + ty * (j.sy1 - j.sy0);
    const float u = px / (float)src.width, v = py / (float)src.height;
    unsigned char *out = const_cast<unsigned char *>(dst.base) + (long long)j.dframe * dst.frame_stride + (long long)y * dst.row_stride + 4ll * x;
    if (u > 1.0f || v > 1.0f || u < 0.0f || v < 0.0f) {
        *reinterpret_cast<unsigned *>(out) = 0u;
        return;
    }
    const float fx = px - 0.5f, fy = py - 0.5f;
    const float x0f = floorf(fx), y0f = floorf(fy);
    const float wx = fx - x0f, wy = fy - y0f;
    const int x0 = (int)x0f, y0 = (int)y0f;
    const int xa = min(max(x0, 0), src.width - 1), xb = min(max(x0 + 1, 0), src.width - 1);
    const int ya = min(max(y0, 0), src.height - 1), yb = min(max(y0 + 1, 0), src.height - 1);
    const unsigned char *sb = src.base + (long long)j.sframe * src.frame_stride;
    const unsigned t00 = *reinterpret_cast<const unsigned *>(sb + (long long)ya * src.row_stride + 4ll * xa);
    const unsigned t01 = *reinterpret_cast<const unsigned *>(sb + (long long)ya * src.row_stride + 4ll * xb);
    const unsigned t10 = *reinterpret_cast<const unsigned *>(sb + (long long)yb * src.row_stride + 4ll * xa);
    const unsigned t11 = *reinterpret_cast<const unsigned *>(sb + (long long)yb * src.row_stride + 4ll * xb);
    unsigned res = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const unsigned a = (t00 >> (8 * c)) & 255u, b = (t01 >> (8 * c)) & 255u, d = (t10 >> (8 * c)) & 255u, e = (t11 >> (8 * c)) & 255u;
        const float fa = c < 3 ? c_srgb_lut[a] : (float)a / 255.0f, fb = c < 3 ? c_srgb_lut[b] : (float)b / 255.0f;
        const float fd = c < 3 ? c_srgb_lut[d] : (float)d / 255.0f, fe = c < 3 ? c_srgb_lut[e] : (float)e / 255.0f;
        const float top = fa + wx * (fb - fa), bot = fd + wx * (fe - fd);
        const float val = top + wy * (bot - top);
        unsigned q;
        if (c < 3) q = srgb_encode_ref(val);
        else q = (unsigned)floor(fmin(fmax((double)val, 0.0), 1.0) * 255.0 + 0.5);
        res |= q << (8 * c);
    }
    *reinterpret_cast<unsigned *>(out) = res;
}

void launch_gather_texels(const FramesDev &f, const ViewDev *views, int n, int out_w, int out_h, uint32_t *out, int max_ctas,
                          cudaStream_t s) {
    g_launch_count++;
    const long long total = (long long)n * out_w * out_h;
    long long ctas = (total + 4 * 256 - 1) / (4 * 256);
    if (ctas > max_ctas) ctas = max_ctas;
    ZB_KNAME("gather_texels_kernel");
    gather_texels_kernel<<<(unsigned)ctas, 256, 0, s>>>(f, views, n, out_w, out_h, out);
}

void launch_view_to_image(const FramesDev &f, const ViewDev *views, int n, int out_w, int out_h, uint8_t *out, cudaStream_t s) {
    g_launch_count++;
    ZB_KNAME("view_to_image_kernel");
    view_to_image_kernel<<<dim3((out_w + 127) / 128, out_h, n), 128, 0, s>>>(f, views, out_w, out_h, reinterpret_cast<unsigned *>(out));
}

void launch_blend(const FramesDev &dst, const FramesDev &src, const void *jobs_dev, int n, int max_w, int max_h, const float *lut_host,
                  cudaStream_t s) {
    g_launch_count++;
    static std::atomic<bool> lut_set[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (!lut_set[dev & 63].load()) {       // sRGB decode table (computed once on the host in f64, like the oracle)
        cudaMemcpyToSymbol(c_srgb_lut, lut_host, 256 * sizeof(float));
        lut_set[dev & 63].store(true);
    }
    ZB_KNAME("blend_kernel");
    blend_kernel<<<dim3((max_w + 31) / 32, (max_h + 7) / 8, n), 256, 0, s>>>(dst, src, reinterpret_cast<const BlendJob *>(jobs_dev));
}

void launch_frames_clear(uint8_t *base, long long frame_stride, long long row_stride, int width, int height, int first,
                         int count, unsigned rgba, cudaStream_t s) {
    g_launch_count++;
    ZB_KNAME("frames_clear_kernel");
    frames_clear_kernel<<<dim3((width + 255) / 256, height, count), 256, 0, s>>>(base, frame_stride, row_stride, width, height, first, rgba);
}

void launch_tracker_prepare(const FramesDev &f, const TrackState *state, int first_frame, int n, int net_w, int net_h,
                            ViewDev *out_views, float *out_fit, ViewHost *out_view_rects, cudaStream_t s) {
    g_launch_count++;
    ZB_KNAME("tracker_prepare_kernel");
    tracker_prepare_kernel<<<(n + 127) / 128, 128, 0, s>>>(f, state, first_frame, n, net_w, net_h, out_views, out_fit,
                                                           out_view_rects);
}

void launch_rois_prepare(const FramesDev &f, const ViewHost *rois, int n, int net_w, int net_h, ViewDev *out_views, float *out_fit,
                         ViewHost *out_view_rects, cudaStream_t s) {
    g_launch_count++;
    ZB_KNAME("rois_prepare_kernel");
    rois_prepare_kernel<<<(n + 127) / 128, 128, 0, s>>>(f, rois, n, net_w, net_h, out_views, out_fit, out_view_rects);
}

void launch_eye_rois(const FramesDev &f, const float *landmarks, const ViewDev *face_views, int n, int num_landmarks, int net_w,
                     int net_h, float margin, ViewDev *out_views, float *out_fit, ViewHost *out_rects, cudaStream_t s) {
    g_launch_count++;
    ZB_KNAME("eye_roi_kernel");
    eye_roi_kernel<<<(2 * n + 127) / 128, 128, 0, s>>>(f, landmarks, face_views, n, num_landmarks, net_w, net_h, margin, out_views,
                                                       out_fit, out_rects);
}

void launch_tracker_update(TrackState *state, const float *out0, int s0, const float *fit, const float *landmarks,
                           const float *scalars, int n, int num_landmarks, float loss_thresh, float roi_padding,
                           int idx_from, int idx_to, float axis_x, float axis_y, ViewHost *out_updated,
                           unsigned char *out_tracked, cudaStream_t s) {
    g_launch_count++;
    ZB_KNAME("tracker_update_kernel");
    tracker_update_kernel<<<n, 128, 0, s>>>(state, out0, s0, fit, landmarks, scalars, num_landmarks, loss_thresh,
                                            roi_padding, idx_from, idx_to, axis_x, axis_y, out_updated, out_tracked);
}

void launch_tracker_set_roi(TrackState *state, const int *ids, const ViewHost *rois, int k, cudaStream_t s) {
    g_launch_count++;
    ZB_KNAME("tracker_set_roi_kernel");
    tracker_set_roi_kernel<<<(k + 127) / 128, 128, 0, s>>>(state, ids, rois, k);
}

void launch_sample(const FramesDev &f, const ViewDev *views, int n, int out_w, int out_h, float lo, float hi,
                   SampleLayout layout, float *out, long long out_img_stride, cudaStream_t s, int round_f16) {
    g_launch_count++;
    dim3 block(128);
    dim3 grid((out_w + 127) / 128, out_h, n);
    ZB_KNAME("sample_kernel");
    sample_kernel<<<grid, block, 0, s>>>(f, views, out_w, out_h, lo, hi, (int)layout, out, out_img_stride, round_f16);
}

// Stem conv (KS x KS, stride 2, Cin = 3 in an NHWC4 tensor or sampled on the fly from `views`).
// Returns false when the layer is outside the kernel's envelope.
bool stem_supported(const ConvDev &p) {
    static const bool disabled = getenv("ZB_NO_STEM") && atoi(getenv("ZB_NO_STEM")) != 0;
    if (disabled) return false;
    if (p.Cs_in != 4 || p.sh != 2 || p.sw != 2 || p.kh != p.kw || (p.kh != 3 && p.kh != 5)) return false;
    if (p.Ns > 64 || p.Ns % 4 || p.Nstore != p.Ns || p.out_pix_stride != p.Ns || p.epi.res || p.epi.act2.kind != ACT_NONE)
        return false;
    if (p.epi.act1.kind == ACT_SIGMOID || p.M % (p.Ho * p.Wo)) return false;
    return true;
}

bool launch_stem(const FramesDev &f, const ViewDev *views, float lo, float hi, const ConvDev &p, cudaStream_t s, int round_f16) {
    if (!stem_supported(p)) return false;
    g_launch_count++;
    // tensor-core stem: sampled views, up to 64 output channels, 8-aligned map width (every bundled network; FaceMeshV2's FLOAT16
    // input included: the rounded mapped value is itself an FP16 number).  ZB_STEM_MMA=0 keeps the SIMT stem for A/Bs
    static const int mma_env = getenv("ZB_STEM_MMA") ? atoi(getenv("ZB_STEM_MMA")) : 15;   // bit 0: 5x5, bit 1: 3x3, bit 2: N = 64, bit 3: FLOAT16 input
    if (views && p.Ns <= 64 && p.Ns >= 8 && p.Wo % 8 == 0 && (mma_env & (p.kh == 5 ? 1 : 2)) && (p.Ns <= 32 || (mma_env & 4)) &&
        (!round_f16 || (mma_env & 8))) {
        static const int minb = getenv("ZB_STEM_MMA_CTAS") ? atoi(getenv("ZB_STEM_MMA_CTAS")) : 5;   // measured: 5 (N = 16) / 4 (N >= 32) CTAs per SM
        bool ok;
#define ZB_STEM_MMA_GO(KS_, N_, MAXB_) \
    (minb >= 5 ? launch_stem_mma_cfg<KS_, N_, MAXB_>(f, views, lo, hi, p, round_f16, s) : minb == 4 ? launch_stem_mma_cfg<KS_, N_, 4>(f, views, lo, hi, p, round_f16, s) \
                                                                                                   : launch_stem_mma_cfg<KS_, N_, 3>(f, views, lo, hi, p, round_f16, s))
        // (a 32-column accumulator read needs > 48 registers: four CTAs per SM at most there)
        if (p.kh == 5) ok = p.Ns <= 16 ? ZB_STEM_MMA_GO(5, 16, 5) : p.Ns <= 32 ? ZB_STEM_MMA_GO(5, 32, 4)
                                                                     : launch_stem_mma_cfg<5, 64, 3>(f, views, lo, hi, p, round_f16, s);   // (40 weight registers: three CTAs)
        else ok = p.Ns <= 16 ? ZB_STEM_MMA_GO(3, 16, 5) : p.Ns <= 32 ? ZB_STEM_MMA_GO(3, 32, 4) : ZB_STEM_MMA_GO(3, 64, 4);
#undef ZB_STEM_MMA_GO
        if (ok) return true;
    }
    // measured: 2 pixels per thread pays for the LDS-bound 5x5 stem (0.79 -> 0.63 ms), not for the 3x3 one
    static const int ppt = getenv("ZB_STEM_PPT") ? atoi(getenv("ZB_STEM_PPT")) : 0;
    if (p.kh == 3) {
        if (p.Ns <= 16) return ppt == 2 ? launch_stem_cfg<3, 16, 2>(f, views, lo, hi, p, round_f16, s) : launch_stem_cfg<3, 16, 1>(f, views, lo, hi, p, round_f16, s);
        if (p.Ns == 24) return ppt == 2 ? launch_stem_cfg<3, 24, 2>(f, views, lo, hi, p, round_f16, s) : launch_stem_cfg<3, 24, 1>(f, views, lo, hi, p, round_f16, s);
        return launch_stem_cfg<3, 32, 1>(f, views, lo, hi, p, round_f16, s);
    }
    if (p.Ns <= 16) return ppt != 1 ? launch_stem_cfg<5, 16, 2>(f, views, lo, hi, p, round_f16, s) : launch_stem_cfg<5, 16, 1>(f, views, lo, hi, p, round_f16, s);
    if (p.Ns == 24) return ppt != 1 ? launch_stem_cfg<5, 24, 2>(f, views, lo, hi, p, round_f16, s) : launch_stem_cfg<5, 24, 1>(f, views, lo, hi, p, round_f16, s);
    return launch_stem_cfg<5, 32, 1>(f, views, lo, hi, p, round_f16, s);
}

void launch_decode_nms(const float *boxes, const float *scores, const float *fit, int n, const DecodeParams &p,
                       DetDev *out, int *out_counts, cudaStream_t s) {
    g_launch_count++;
    const size_t smem = (size_t)p.num_anchors * (sizeof(float4) + sizeof(float) + 4 * sizeof(int));
    static SmemOptIn opt_in;
    opt_in.ensure(decode_nms_kernel, smem);   // a failure surfaces as the launch error
    ZB_KNAME("decode_nms_kernel");
    decode_nms_kernel<<<n, 256, smem, s>>>(boxes, scores, fit, p, out, out_counts);
}

void launch_face_roi(const FramesDev &f, const DetDev *dets, const int *counts, int cap, int first_frame, int n,
                     int net_w, int net_h, ViewDev *out_views, float *out_fit, ViewHost *out_view_rects,
                     cudaStream_t s, float grow_rel_amount, int use_angle) {
    g_launch_count++;
    ZB_KNAME("face_roi_kernel");
    face_roi_kernel<<<(n + 127) / 128, 128, 0, s>>>(f, dets, counts, cap, first_frame, n, net_w, net_h, out_views,
                                                    out_fit, out_view_rects, grow_rel_amount, use_angle);
}

void launch_landmarks(const float *out0, int s0, const float *out1, int s1, const float *out2, int s2,
                      const float *fit, const ViewDev *views, const ViewHost *view_rects, int n,
                      const LandmarkParams &p, float *landmarks, float *scalars, cudaStream_t s, const FilterDev *filter,
                      const int *sel) {
    g_launch_count++;
    dim3 grid((unsigned)((p.num_landmarks + 127) / 128) * (unsigned)n);
    FilterDev none{};
    ZB_KNAME("landmarks_kernel");
    landmarks_kernel<<<grid, 128, 0, s>>>(out0, s0, out1, s1, out2, s2, fit, views, view_rects, n, p, landmarks,
                                          scalars, filter ? *filter : none, sel);
}

void launch_compact_views(const ViewDev *views, int n, ViewDev *out_views, int *sel, int *count, float *scalars, cudaStream_t s) {
    g_launch_count++;
    ZB_KNAME("compact_views_kernel");
    compact_views_kernel<<<1, 1024, 0, s>>>(views, n, out_views, sel, count, scalars);
}

void launch_filter_apply(const FilterDev &f, float *values, long long count, cudaStream_t s) {
    g_launch_count++;
    ZB_KNAME("filter_apply_kernel");
    filter_apply_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(f, values, count);
}

}  // namespace zb
