// sm_100a thin fused block kernel (see the comment on dwpw_thin_kernel).  Split from kernels_conv.cu so the
// two translation units compile in parallel.
#include <cuda_runtime.h>

#include <cstdlib>
#include <type_traits>

#include "conv_common.cuh"
#include "kernels.h"

namespace zb {
namespace {

// ------------------------------------------------------------------------------------------------
// Thin fused block: depthwise 3x3 -> (mid act) -> pointwise 1x1 -> bias -> act1 -> +residual -> act2 for
// layers with few channels on large maps (Cs_in <= 48, Cout <= 64: the first blocks of every network).
// One CTA = one TW x TH output tile of one image, one thread = one output pixel:
//   * the input halo tile is staged ONCE in shared memory with coalesced 128-bit loads (pixel stride padded
//     to CS+4 words so the per-pixel 128-bit reads below are bank-conflict free);
//   * the depthwise result (CS values) and the pointwise accumulators (NP values) live in registers;
//   * depthwise / pointwise weights, bias and PReLU slopes are read from shared memory as warp broadcasts;
//   * the residual (same tensor as the input in every Blaze-style block, optionally through the 2x2 max-pool
//     of stride-2 blocks) is taken from the staged tile, so each input element is read from HBM/L2 once.
// ------------------------------------------------------------------------------------------------
template <int CS, int S, int NP, int TW, int TH>
__global__ void __launch_bounds__(TW *TH) dwpw_thin_kernel(const ConvDev p, int tiles_x, int tiles_y, int NSP) {
    constexpr int NT = TW * TH;
    constexpr int PS = CS + 4;                       // padded pixel stride in words
    constexpr int IW = (TW - 1) * S + 3, IH = (TH - 1) * S + 3;
    constexpr int CQ = CS / 4;
    extern __shared__ __align__(16) float smem[];
    float *s_in = smem;                              // [IH][IW][PS]
    float *s_dww = s_in + IH * IW * PS;              // [9][CS]
    float *s_dwb = s_dww + 9 * CS;                   // [CS]
    float *s_pw = s_dwb + CS;                        // [CS][NSP]
    float *s_pb = s_pw + CS * NSP;                   // [NSP]
    float *s_sl = s_pb + NSP;                        // [NSP] act2 PReLU slopes (if any)

    const int tid = threadIdx.x;
    int b = blockIdx.x;
    const int tile_x = b % tiles_x;
    b /= tiles_x;
    const int tile_y = b % tiles_y;
    const int img = b / tiles_y;
    const int oy0 = tile_y * TH, ox0 = tile_x * TW;
    const int iy_org = oy0 * S - p.pt, ix_org = ox0 * S - p.pl;
    const float *in_img = p.in + (long long)img * p.in_img_stride;

    // --- stage input halo tile + weights with cp.async (LDGSTS): every copy of the CTA is in flight at once
    // instead of one dependent LDG->STS round trip per loop iteration; out-of-image pixels are zero-filled.
    // row by row, per-thread (column, quad) arithmetic hoisted out of the row loop
    constexpr int ROW_CHUNKS = IW * CQ, CPT = (ROW_CHUNKS + NT - 1) / NT;
#pragma unroll
    for (int u = 0; u < CPT; u++) {
        const int c = tid + u * NT;
        if (c < ROW_CHUNKS) {
            const int tx = c / CQ, q = c - tx * CQ;
            const int ix = ix_org + tx;
            const bool col_ok = ix >= 0 && ix < p.W;
            float *dst = s_in + tx * PS + q * 4;
            const float *src = in_img + ((long long)iy_org * p.W + ix) * CS + q * 4;
#pragma unroll 2
            for (int ty = 0; ty < IH; ty++) {
                const int iy = iy_org + ty;
                const bool ok = col_ok && iy >= 0 && iy < p.H;
                cp_async16(dst, ok ? src : in_img, ok ? 16 : 0);
                dst += IW * PS;
                src += (long long)p.W * CS;
            }
        }
    }
    for (int e = tid; e < 9 * CS / 4; e += NT) cp_async16(s_dww + e * 4, p.dw_w + e * 4, 16);
    for (int e = tid; e < CS / 4; e += NT) cp_async16(s_dwb + e * 4, p.dw_b + e * 4, 16);
    for (int e = tid; e < CS * NSP / 4; e += NT) {
        const int k = e / (NSP / 4), nq = e - k * (NSP / 4);
        const bool ok = nq * 4 < p.Ns;
        cp_async16(s_pw + e * 4, ok ? p.w + (long long)k * p.Ns + nq * 4 : p.w, ok ? 16 : 0);
    }
    for (int e = tid; e < NSP; e += NT) {
        s_pb[e] = e < p.Ns ? __ldg(p.epi.bias + e) : 0.f;
        s_sl[e] = (p.epi.act2.kind == ACT_PRELU && e < p.Ns) ? __ldg(p.epi.act2.slope + e) : 0.f;
    }
    cp_async_wait_all();
    __syncthreads();

    const int tx = tid % TW, ty = tid / TW;
    const int oy = oy0 + ty, ox = ox0 + tx;
    if (oy >= p.Ho || ox >= p.Wo) return;

    // --- depthwise 3x3 into registers ------------------------------------------------------------------
    float dwo[CS];
    const float *s_px = s_in + ((ty * S) * IW + tx * S) * PS;
#pragma unroll
    for (int q = 0; q < CQ; q++) {
        float4 v = *reinterpret_cast<const float4 *>(s_dwb + q * 4);
#pragma unroll
        for (int t = 0; t < 9; t++) {
            const float4 x = *reinterpret_cast<const float4 *>(s_px + ((t / 3) * IW + (t % 3)) * PS + q * 4);
            const float4 wv = *reinterpret_cast<const float4 *>(s_dww + t * CS + q * 4);
            fma4(v, x, wv);
        }
        act4(v, p.act_mid, q * 4);
        dwo[q * 4 + 0] = v.x, dwo[q * 4 + 1] = v.y, dwo[q * 4 + 2] = v.z, dwo[q * 4 + 3] = v.w;
    }

    // --- pointwise in passes of NP output channels --------------------------------------------------------
    const EpiDev &e = p.epi;
    const bool res_smem = e.res == p.in;               // Blaze-style skip: residual is the block input
    float *orow = p.out + (long long)img * p.out_img_stride + ((long long)oy * p.Wo + ox) * p.out_pix_stride;
    for (int n0 = 0; n0 < p.Nstore; n0 += NP) {
        float2 acc[NP / 2];
#pragma unroll
        for (int j = 0; j < NP; j += 4) {
            const float4 bv = *reinterpret_cast<const float4 *>(s_pb + n0 + j);
            acc[j / 2] = make_float2(bv.x, bv.y), acc[j / 2 + 1] = make_float2(bv.z, bv.w);
        }
#pragma unroll
        for (int k = 0; k < CS; k++) {
            const float a = dwo[k];
#pragma unroll
            for (int j = 0; j < NP; j += 4) {
                const float4 wv = *reinterpret_cast<const float4 *>(s_pw + k * NSP + n0 + j);
                fma4s(acc[j / 2], acc[j / 2 + 1], a, wv);
            }
        }
#pragma unroll
        for (int j = 0; j < NP; j += 4) {
            const int n = n0 + j;
            if (n >= p.Nstore) break;
            float v[4] = {acc[j / 2].x, acc[j / 2].y, acc[j / 2 + 1].x, acc[j / 2 + 1].y};
            act4(v, e.act1, n);
            if (e.res) {
                float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
                if (res_smem) {
                    if (n < CS) {
                        if (!e.res_pool) {
                            // stride-1 block: centre pixel of the 3x3 window (pads 1,1)
                            rr = *reinterpret_cast<const float4 *>(s_px + (p.pt * IW + p.pl) * PS + n);
                        } else {
                            // stride-2 block: 2x2 max-pool of input rows/cols (2oy, 2oy+1) x (2ox, 2ox+1)
                            const float *r0 = s_in + ((2 * ty + p.pt) * IW + 2 * tx + p.pl) * PS + n;
                            const float4 a0 = *reinterpret_cast<const float4 *>(r0);
                            const float4 a1 = *reinterpret_cast<const float4 *>(r0 + PS);
                            const float4 a2 = *reinterpret_cast<const float4 *>(r0 + IW * PS);
                            const float4 a3 = *reinterpret_cast<const float4 *>(r0 + IW * PS + PS);
                            rr = make_float4(fmaxf(fmaxf(a0.x, a1.x), fmaxf(a2.x, a3.x)), fmaxf(fmaxf(a0.y, a1.y), fmaxf(a2.y, a3.y)),
                                             fmaxf(fmaxf(a0.z, a1.z), fmaxf(a2.z, a3.z)), fmaxf(fmaxf(a0.w, a1.w), fmaxf(a2.w, a3.w)));
                        }
                    }
                } else {
                    rr = residual4_at(e, img, oy, ox, n);
                }
                v[0] += rr.x, v[1] += rr.y, v[2] += rr.z, v[3] += rr.w;
            }
            if (e.act2.kind == ACT_PRELU) {
                const float4 sl = *reinterpret_cast<const float4 *>(s_sl + n);
                v[0] = v[0] < 0.f ? v[0] * sl.x : v[0];
                v[1] = v[1] < 0.f ? v[1] * sl.y : v[1];
                v[2] = v[2] < 0.f ? v[2] * sl.z : v[2];
                v[3] = v[3] < 0.f ? v[3] * sl.w : v[3];
            } else {
                act4(v, e.act2, n);
            }
            *reinterpret_cast<float4 *>(orow + n) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Register-blocked variant of the thin block ("strip"): one thread = PXV vertically adjacent output pixels.
// ncu on dwpw_thin_kernel<16,1,16> at the bench batch showed the L1/shared pipe 95 % busy (every pixel re-read
// its 9 taps AND all depthwise + pointwise weights from shared memory: 140 LDS.128 for 400 FMA).  Here
//   * the (PXV-1)*S+3 input rows of a strip are loaded once and each feeds up to 3 output pixels,
//   * every weight quad fetched from shared memory is used for PXV pixels,
// which cuts shared-memory wavefronts per pixel ~3x at PXV = 4.  Lanes run along x with pixel stride CS+4 words
// (conflict-free 128-bit reads); stride-2 blocks stage even and odd input columns in separate planes so the
// lane stride stays CS+4 words.  FMA order per output value is the same as in every other kernel of the
// library (bias, taps row-major; bias, k ascending), so results are bit-identical to them.
// ------------------------------------------------------------------------------------------------
template <int CS, int S, int NP, int PXV, int TW, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) dwpw_strip_kernel(const ConvDev p, int tiles_x, int tiles_y) {
    constexpr int NT = 32 * WARPS;
    constexpr int SUB = 32 / TW;                       // pixel strips per warp along y
    constexpr int TH = WARPS * SUB * PXV;              // output rows per CTA
    constexpr int PS = CS + 4, CQ = CS / 4;
    constexpr int IH = (TH - 1) * S + 3;
    constexpr int IW = (TW - 1) * S + 3;               // input columns needed
    constexpr int PW = S == 1 ? IW : TW + 1;           // columns per plane (S == 2: even plane TW+1, odd plane TW)
    constexpr int PLANES = S;
    constexpr int ROWS = (PXV - 1) * S + 3;
    extern __shared__ __align__(16) float smem[];
    float *s_in = smem;                                // [IH][PLANES][PW][PS]
    float *s_dww = s_in + IH * PLANES * PW * PS;       // [9][CS]
    float *s_dwb = s_dww + 9 * CS;                     // [CS]
    float *s_pw = s_dwb + CS;                          // [CS][NP]
    float *s_pb = s_pw + CS * NP;                      // [NP]
    float *s_sl = s_pb + NP;                           // [NP]

    const int tid = threadIdx.x;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y, img = blockIdx.z;
    const int oy0 = tile_y * TH, ox0 = tile_x * TW;
    const int iy_org = oy0 * S - p.pt, ix_org = ox0 * S - p.pl;
    const float *in_img = p.in + (long long)img * p.in_img_stride;
    pdl_trigger();
    pdl_wait();                                        // the previous kernel's activations are complete and visible

    // Halo fill, row by row: a thread's (column, channel quad) pairs are fixed, so the index arithmetic (the
    // divisions cost a quarter of the kernel's instructions when done per 16-byte chunk) is hoisted out of the row loop.
    constexpr int ROW_CHUNKS = IW * CQ;                // 16-byte chunks per staged input row
    constexpr int CPT = (ROW_CHUNKS + NT - 1) / NT;    // chunks per thread per row
#pragma unroll
    for (int u = 0; u < CPT; u++) {
        const int c = tid + u * NT;
        if (c < ROW_CHUNKS) {
            const int tx = c / CQ, q = c - tx * CQ;
            const int ix = ix_org + tx;
            const bool col_ok = ix >= 0 && ix < p.W;
            const int slot = S == 1 ? tx : ((tx & 1) * PW + (tx >> 1));
            float *dst = s_in + slot * PS + q * 4;
            const float *src = in_img + ((long long)iy_org * p.W + ix) * CS + q * 4;
#pragma unroll 2
            for (int ty = 0; ty < IH; ty++) {
                const int iy = iy_org + ty;
                const bool ok = col_ok && iy >= 0 && iy < p.H;
                cp_async16(dst, ok ? src : in_img, ok ? 16 : 0);
                dst += PLANES * PW * PS;
                src += (long long)p.W * CS;
            }
        }
    }
    for (int e = tid; e < 9 * CQ; e += NT) cp_async16(s_dww + e * 4, p.dw_w + e * 4, 16);
    for (int e = tid; e < CQ; e += NT) cp_async16(s_dwb + e * 4, p.dw_b + e * 4, 16);
    for (int e = tid; e < CS * NP / 4; e += NT) {
        const int k = e / (NP / 4), nq = e - k * (NP / 4);
        const bool ok = nq * 4 < p.Ns;
        cp_async16(s_pw + e * 4, ok ? p.w + (long long)k * p.Ns + nq * 4 : p.w, ok ? 16 : 0);
    }
    for (int e = tid; e < NP; e += NT) {
        s_pb[e] = e < p.Ns ? __ldg(p.epi.bias + e) : 0.f;
        s_sl[e] = (p.epi.act2.kind == ACT_PRELU && e < p.Ns) ? __ldg(p.epi.act2.slope + e) : 0.f;
    }
    cp_async_wait_all();
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    const int tx = lane % TW;
    const int ty0 = (warp * SUB + lane / TW) * PXV;    // first output row of this thread's strip (tile-local)
    const int ox = ox0 + tx;
    if (ox >= p.Wo || oy0 + ty0 >= p.Ho) return;

    // tap (r, kx) of this thread: input row ty0*S + r, input column tx*S + kx
    auto tap = [&](int r, int kx) -> const float * {
        const int row = ty0 * S + r;
        const int slot = S == 1 ? (row * PW + tx + kx) : ((row * 2 + (kx & 1)) * PW + tx + (kx >> 1));
        return s_in + slot * PS;
    };

    float2 acc[PXV][NP / 2];
#pragma unroll
    for (int j = 0; j < NP; j += 4) {
        const float4 bv = *reinterpret_cast<const float4 *>(s_pb + j);
#pragma unroll
        for (int i = 0; i < PXV; i++) acc[i][j / 2] = make_float2(bv.x, bv.y), acc[i][j / 2 + 1] = make_float2(bv.z, bv.w);
    }

    const bool has_mid = p.act_mid.kind != ACT_NONE;
#pragma unroll 1
    for (int q = 0; q < CQ; q++) {
        float4 w9[9];
#pragma unroll
        for (int t = 0; t < 9; t++) w9[t] = *reinterpret_cast<const float4 *>(s_dww + t * CS + q * 4);
        const float4 bias = *reinterpret_cast<const float4 *>(s_dwb + q * 4);
        float4 v[PXV];
#pragma unroll
        for (int i = 0; i < PXV; i++) v[i] = bias;
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
            const float4 x0 = *reinterpret_cast<const float4 *>(tap(r, 0) + q * 4);
            const float4 x1 = *reinterpret_cast<const float4 *>(tap(r, 1) + q * 4);
            const float4 x2 = *reinterpret_cast<const float4 *>(tap(r, 2) + q * 4);
#pragma unroll
            for (int i = 0; i < PXV; i++) {
                const int ky = r - i * S;
                if (ky >= 0 && ky < 3) {
                    fma4(v[i], x0, w9[ky * 3 + 0]);
                    fma4(v[i], x1, w9[ky * 3 + 1]);
                    fma4(v[i], x2, w9[ky * 3 + 2]);
                }
            }
        }
        if (has_mid) {
#pragma unroll
            for (int i = 0; i < PXV; i++) act4(v[i], p.act_mid, q * 4);
        }
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            const float *wrow = s_pw + (q * 4 + kk) * NP;
#pragma unroll
            for (int j = 0; j < NP; j += 4) {
                const float4 wv = *reinterpret_cast<const float4 *>(wrow + j);
#pragma unroll
                for (int i = 0; i < PXV; i++) {
                    const float a = kk == 0 ? v[i].x : kk == 1 ? v[i].y : kk == 2 ? v[i].z : v[i].w;
                    fma4s(acc[i][j / 2], acc[i][j / 2 + 1], a, wv);
                }
            }
        }
    }

    const EpiDev &e = p.epi;
    const bool res_smem = e.res == p.in;
    // MODE 0: generic epilogue.  MODE 1 / 2: the Blaze block (no act1, residual = this block's input taken from the
    // staged tile, then ReLU / PReLU, every channel group stored) with all the uniform branches resolved at compile
    // time - they were ~18 % of the kernel's instructions.
    auto epilogue = [&](auto mode_tag) {
        constexpr int MODE = decltype(mode_tag)::value;
#pragma unroll
        for (int i = 0; i < PXV; i++) {
            const int oy = oy0 + ty0 + i;
            if (oy >= p.Ho) break;
            float *orow = p.out + (long long)img * p.out_img_stride + ((long long)oy * p.Wo + ox) * p.out_pix_stride;
            // Blaze fast path, stride 1: all residual quads of this pixel are fetched from the staged tile before any
            // is consumed (independent LDS in flight instead of LDS -> FADD -> compare chains: 14 % of the stall samples)
            float4 rpre[NP / 4];
            if (MODE != 0 && S == 1) {
#pragma unroll
                for (int j = 0; j < NP; j += 4)
                    rpre[j / 4] = j < CS ? *reinterpret_cast<const float4 *>(tap(i * S + p.pt, p.pl) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < NP; j += 4) {
                if (MODE == 0 && j >= p.Nstore) break;
                float v4[4] = {acc[i][j / 2].x, acc[i][j / 2].y, acc[i][j / 2 + 1].x, acc[i][j / 2 + 1].y};
                if (MODE == 0) act4(v4, e.act1, j);
                if (MODE != 0 && S == 1) {
                    const float4 rr = rpre[j / 4];
                    v4[0] += rr.x, v4[1] += rr.y, v4[2] += rr.z, v4[3] += rr.w;
                } else if (MODE != 0 || e.res) {
                    float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (MODE != 0 || res_smem) {
                        if (j < CS) {
                            if (!e.res_pool) {
                                rr = *reinterpret_cast<const float4 *>(tap(i * S + p.pt, p.pl) + j);
                            } else {
                                const float4 a0 = *reinterpret_cast<const float4 *>(tap(i * S + p.pt, p.pl) + j);
                                const float4 a1 = *reinterpret_cast<const float4 *>(tap(i * S + p.pt, p.pl + 1) + j);
                                const float4 a2 = *reinterpret_cast<const float4 *>(tap(i * S + p.pt + 1, p.pl) + j);
                                const float4 a3 = *reinterpret_cast<const float4 *>(tap(i * S + p.pt + 1, p.pl + 1) + j);
                                rr = make_float4(fmaxf(fmaxf(a0.x, a1.x), fmaxf(a2.x, a3.x)), fmaxf(fmaxf(a0.y, a1.y), fmaxf(a2.y, a3.y)),
                                                 fmaxf(fmaxf(a0.z, a1.z), fmaxf(a2.z, a3.z)), fmaxf(fmaxf(a0.w, a1.w), fmaxf(a2.w, a3.w)));
                            }
                        }
                    } else {
                        rr = residual4_at(e, img, oy, ox, j);
                    }
                    v4[0] += rr.x, v4[1] += rr.y, v4[2] += rr.z, v4[3] += rr.w;
                }
                if (MODE == 2 || (MODE == 0 && e.act2.kind == ACT_PRELU)) {
                    const float4 sl = *reinterpret_cast<const float4 *>(s_sl + j);
                    v4[0] = v4[0] < 0.f ? v4[0] * sl.x : v4[0];
                    v4[1] = v4[1] < 0.f ? v4[1] * sl.y : v4[1];
                    v4[2] = v4[2] < 0.f ? v4[2] * sl.z : v4[2];
                    v4[3] = v4[3] < 0.f ? v4[3] * sl.w : v4[3];
                } else if (MODE == 1) {
#pragma unroll
                    for (int c = 0; c < 4; c++) v4[c] = fmaxf(v4[c], 0.0f);
                } else {
                    act4(v4, e.act2, j);
                }
                *reinterpret_cast<float4 *>(orow + j) = make_float4(v4[0], v4[1], v4[2], v4[3]);
            }
        }
    };
    const bool blaze = e.act1.kind == ACT_NONE && e.res && res_smem && p.Nstore >= NP && (S == 2 || !e.res_pool) &&
                       (e.act2.kind == ACT_RELU || e.act2.kind == ACT_PRELU);
    if (!blaze) epilogue(std::integral_constant<int, 0>{});
    else if (e.act2.kind == ACT_RELU) epilogue(std::integral_constant<int, 1>{});
    else epilogue(std::integral_constant<int, 2>{});
}

template <int CS, int S, int NP, int PXV, int TW, int WARPS>
bool launch_strip_cfg(const ConvDev &p, cudaStream_t s) {
    constexpr int SUB = 32 / TW, TH = WARPS * SUB * PXV;
    constexpr int IH = (TH - 1) * S + 3, IW = (TW - 1) * S + 3, PW = S == 1 ? IW : TW + 1;
    const size_t smem = sizeof(float) * ((size_t)IH * S * PW * (CS + 4) + 10 * CS + (size_t)CS * NP + 2 * NP);
    auto kern = dwpw_strip_kernel<CS, S, NP, PXV, TW, WARPS>;
    static SmemOptIn opt_in;
    if (!opt_in.ensure(kern, smem)) return false;
    const int tiles_x = (p.Wo + TW - 1) / TW, tiles_y = (p.Ho + TH - 1) / TH;
    const int images = p.M / (p.Ho * p.Wo);
    ZB_KNAME("dwpw_strip_kernel", CS, S, NP, PXV, TW, WARPS);
    launch_pdl(4, kern, dim3(tiles_x, tiles_y, images), dim3(32 * WARPS), smem, s, p, tiles_x, tiles_y);
    return true;
}

// Envelope of the strip kernel: Cs_in 8, 16 or 24, all output channels in one register pass (PXV * Ns <= 64).
bool launch_dwpw_strip(const ConvDev &p, cudaStream_t s) {
    static const bool disabled = getenv("ZB_NO_STRIP") && atoi(getenv("ZB_NO_STRIP")) != 0;
    if (disabled || p.Ns > 32) return false;
    const bool wide = p.Wo % 32 == 0 || p.Wo > 64;       // 32-pixel-wide tiles unless that wastes half a tile
    if (p.Cs_in == 8) {   // FaceMeshV2 / full-range BlazeFace first blocks (128x128x8 -> 16, 96x96x8 -> 32)
        if (p.sh != 1) return false;
        if (p.Ns <= 16) return wide ? launch_strip_cfg<8, 1, 16, 4, 32, 4>(p, s) : launch_strip_cfg<8, 1, 16, 4, 16, 2>(p, s);
        return wide ? launch_strip_cfg<8, 1, 32, 2, 32, 4>(p, s) : launch_strip_cfg<8, 1, 32, 2, 16, 2>(p, s);
    }
    if (p.Cs_in == 16) {
        if (p.sh == 1) {
            if (p.Ns <= 16) return wide ? launch_strip_cfg<16, 1, 16, 4, 32, 4>(p, s) : launch_strip_cfg<16, 1, 16, 4, 16, 2>(p, s);
            return wide ? launch_strip_cfg<16, 1, 32, 2, 32, 4>(p, s) : launch_strip_cfg<16, 1, 32, 2, 16, 2>(p, s);
        }
        if (p.Ns <= 16) return launch_strip_cfg<16, 2, 16, 4, 16, 2>(p, s);
        return launch_strip_cfg<16, 2, 32, 2, 16, 2>(p, s);
    }
    if (p.Cs_in == 24) {
        if (p.sh == 1) {
            if (p.Ns <= 24) return wide ? launch_strip_cfg<24, 1, 24, 2, 32, 4>(p, s) : launch_strip_cfg<24, 1, 24, 2, 16, 2>(p, s);
            return wide ? launch_strip_cfg<24, 1, 32, 2, 32, 4>(p, s) : launch_strip_cfg<24, 1, 32, 2, 16, 2>(p, s);
        }
        if (p.Ns <= 24) return launch_strip_cfg<24, 2, 24, 2, 16, 2>(p, s);
        return launch_strip_cfg<24, 2, 32, 2, 16, 2>(p, s);
    }
    return false;
}

template <int CS, int S, int NP>
bool launch_thin_cfg(const ConvDev &p, cudaStream_t s) {
    constexpr int TW = S == 1 ? 32 : 16, TH = 8;
    constexpr int IW = (TW - 1) * S + 3, IH = (TH - 1) * S + 3;
    const int NSP = (p.Ns + NP - 1) / NP * NP;
    const size_t smem = sizeof(float) * ((size_t)IH * IW * (CS + 4) + 9 * CS + CS + (size_t)CS * NSP + 2 * NSP);
    if (smem > 200 * 1024) return false;
    auto kern = dwpw_thin_kernel<CS, S, NP, TW, TH>;
    static SmemOptIn opt_in;
    if (!opt_in.ensure(kern, smem)) return false;
    const int tiles_x = (p.Wo + TW - 1) / TW, tiles_y = (p.Ho + TH - 1) / TH;
    const int images = p.M / (p.Ho * p.Wo);
    ZB_KNAME("dwpw_thin_kernel", CS, S, NP, TW, TH);
    kern<<<(unsigned)(tiles_x * tiles_y * images), TW * TH, smem, s>>>(p, tiles_x, tiles_y, NSP);
    return true;
}

template <int CS>
bool launch_thin_cs(const ConvDev &p, cudaStream_t s) {
    if (p.sh == 1) return p.Ns <= 16 ? launch_thin_cfg<CS, 1, 16>(p, s) : launch_thin_cfg<CS, 1, 32>(p, s);
    return p.Ns <= 16 ? launch_thin_cfg<CS, 2, 16>(p, s) : launch_thin_cfg<CS, 2, 32>(p, s);
}

// ------------------------------------------------------------------------------------------------
// Thin pointwise conv (1x1, stride 1) for few channels on large maps - the projection convs of inverted-residual
// blocks (FaceMeshV2: 128x128x16 -> 8, 64x64x32 -> 16): one thread = PX pixels (128 pixels apart, so a warp's loads
// are contiguous), the K x N weights are read from shared memory as broadcasts once per PX pixels, packed FFMA2.
// The GEMM tile spends 0.47 ms on 512 x 128x128x16 -> 8 (memory floor 0.12 ms): K = 16 is one half-empty k-step.
// ------------------------------------------------------------------------------------------------
template <int CIN, int NP, int PX>
__global__ void __launch_bounds__(128) pw_thin_kernel(const ConvDev p) {
    __shared__ __align__(16) float s_w[CIN * NP];
    __shared__ __align__(16) float s_b[NP], s_s1[NP], s_s2[NP];
    const int tid = threadIdx.x;
    for (int e = tid; e < CIN * NP; e += 128) {
        const int k = e / NP, n = e - k * NP;
        s_w[e] = n < p.Ns ? __ldg(p.w + (long long)k * p.Ns + n) : 0.f;
    }
    const EpiDev &e = p.epi;
    for (int n = tid; n < NP; n += 128) {
        s_b[n] = n < p.Ns ? __ldg(e.bias + n) : 0.f;
        s_s1[n] = (e.act1.kind == ACT_PRELU && n < p.Ns) ? __ldg(e.act1.slope + n) : 0.f;
        s_s2[n] = (e.act2.kind == ACT_PRELU && n < p.Ns) ? __ldg(e.act2.slope + n) : 0.f;
    }
    __syncthreads();
    const long long m0 = (long long)blockIdx.x * (128 * PX) + tid;
    const int HoWo = p.Ho * p.Wo;
    float4 x[PX][CIN / 4];
#pragma unroll
    for (int i = 0; i < PX; i++) {
        const long long m = m0 + 128 * i;
        if (m < p.M) {
            const int img = (int)(m / HoWo);
            const float *src = p.in + (long long)img * p.in_img_stride + (m - (long long)img * HoWo) * CIN;
#pragma unroll
            for (int q = 0; q < CIN / 4; q++) x[i][q] = ldg4(src + q * 4);
        } else {
#pragma unroll
            for (int q = 0; q < CIN / 4; q++) x[i][q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    float2 acc[PX][NP / 2];
#pragma unroll
    for (int j = 0; j < NP; j += 2)
#pragma unroll
        for (int i = 0; i < PX; i++) acc[i][j / 2] = make_float2(s_b[j], s_b[j + 1]);
#pragma unroll
    for (int k = 0; k < CIN; k++) {
#pragma unroll
        for (int j = 0; j < NP; j += 4) {
            const float4 wv = *reinterpret_cast<const float4 *>(s_w + k * NP + j);
#pragma unroll
            for (int i = 0; i < PX; i++) {
                const float4 xq = x[i][k / 4];
                const float a = (k & 3) == 0 ? xq.x : (k & 3) == 1 ? xq.y : (k & 3) == 2 ? xq.z : xq.w;
                fma4s(acc[i][j / 2], acc[i][j / 2 + 1], a, wv);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < PX; i++) {
        const long long m = m0 + 128 * i;
        if (m >= p.M) continue;
        const int img = (int)(m / HoWo);
        const int r = (int)(m - (long long)img * HoWo);
        const int oy = r / p.Wo, ox = r - oy * p.Wo;
        float *orow = p.out + (long long)img * p.out_img_stride + (long long)r * p.out_pix_stride;
#pragma unroll
        for (int j = 0; j < NP; j += 4) {
            if (j >= p.Nstore) break;
            float v[4] = {acc[i][j / 2].x, acc[i][j / 2].y, acc[i][j / 2 + 1].x, acc[i][j / 2 + 1].y};
            if (e.act1.kind == ACT_PRELU) {
#pragma unroll
                for (int c = 0; c < 4; c++) v[c] = v[c] < 0.f ? v[c] * s_s1[j + c] : v[c];
            } else {
                act4(v, e.act1, j);
            }
            if (e.res) {
                const float4 rr = residual4_at(e, img, oy, ox, j);
                v[0] += rr.x, v[1] += rr.y, v[2] += rr.z, v[3] += rr.w;
            }
            if (e.act2.kind == ACT_PRELU) {
#pragma unroll
                for (int c = 0; c < 4; c++) v[c] = v[c] < 0.f ? v[c] * s_s2[j + c] : v[c];
            } else {
                act4(v, e.act2, j);
            }
            *reinterpret_cast<float4 *>(orow + j) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Dense heads: a convolution whose window IS the whole input map (Ho = Wo = 1: the Gemm / "conv kh x kw over a kh x kw map"
// heads of the landmark networks, M = images rows of K = kh * kw * Cs_in contiguous inputs, N = 1 ... a few hundred).  The
// tcgen05 GEMM runs such a layer as M / 128 CTAs walking K in 32-wide chunks - two CTAs and 21 serial chunks for the hand
// network's 672 -> 63 head: 35 us for 11 MFLOP.  Here a CTA owns R rows and 32 output columns (lane = column: the weight row
// is one coalesced 128-byte read shared by the R rows, the inputs are broadcast float4 reads), f32 FMA.
// ------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256) dense_head_kernel(const ConvDev p) {
    // a CTA = R rows x 32 output columns; its eight warps split K (the layer is a latency chain of L2 round trips otherwise:
    // the first version, one warp per R rows walking all of K, took 140 us for the 672 -> 63 head) and reduce through shared memory
    __shared__ float s_part[8][R][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * 32 + lane;
    const int m0 = blockIdx.y * R;
    const float *w = p.w + (n < p.Ns ? n : 0);
    pdl_trigger();
    pdl_wait();                                                       // the previous kernel's activations are complete and visible
    const float *in[R];
#pragma unroll
    for (int r = 0; r < R; r++) in[r] = p.in + (long long)min(m0 + r, p.M - 1) * p.in_img_stride;
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = 0.0f;
    const int kc = ((p.K / 4 + 7) / 8) * 4;                          // K slice of a warp (multiple of 4)
    const int k_end = min(p.K, (warp + 1) * kc);
#pragma unroll 4
    for (int k = warp * kc; k < k_end; k += 4) {
        const float w0 = __ldg(w + (long long)k * p.Ns), w1 = __ldg(w + (long long)(k + 1) * p.Ns),
                    w2 = __ldg(w + (long long)(k + 2) * p.Ns), w3 = __ldg(w + (long long)(k + 3) * p.Ns);
#pragma unroll
        for (int r = 0; r < R; r++) {
            const float4 x = ldg4(in[r] + k);
            acc[r] = fmaf(x.x, w0, acc[r]);
            acc[r] = fmaf(x.y, w1, acc[r]);
            acc[r] = fmaf(x.z, w2, acc[r]);
            acc[r] = fmaf(x.w, w3, acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++) s_part[warp][r][lane] = acc[r];
    __syncthreads();
    if (warp >= R || m0 + warp >= p.M || n >= p.Nstore) return;       // warp r finishes row r
    float v = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) v += s_part[i][warp][lane];
    v = n < p.Ns ? v + __ldg(p.epi.bias + n) : 0.0f;
    v = apply_act(v, p.epi.act1, n);
    v = apply_act(v, p.epi.act2, n);
    p.out[(long long)(m0 + warp) * p.out_img_stride + n] = v;
}

template <int CIN, int NP, int PX>
bool launch_pw_thin_cfg(const ConvDev &p, cudaStream_t s) {
    ZB_KNAME("pw_thin_kernel", CIN, NP, PX);
    pw_thin_kernel<CIN, NP, PX><<<(unsigned)((p.M + 128 * PX - 1) / (128 * PX)), 128, 0, s>>>(p);
    return true;
}

}  // namespace

// Returns false when the layer is outside the thin kernel's envelope (the GEMM-tile kernel handles it).
bool dwpw_thin_supported(const ConvDev &p) {
    static const bool disabled = getenv("ZB_NO_THIN") && atoi(getenv("ZB_NO_THIN")) != 0;
    if (disabled) return false;
    if (p.kh != 3 || p.kw != 3 || p.sh != p.sw || (p.sh != 1 && p.sh != 2)) return false;
    if (p.Cs_in > 48 || p.Ns > 64 || p.Ns % 4 || p.Nstore != p.Ns || p.out_pix_stride != p.Ns) return false;
    if (p.Ho * p.Wo < 1024) return false;            // small maps: tile quantisation wastes the CTA
    if (p.sh == 1 && !(p.pt == 1 && p.pl == 1)) return false;
    if (p.sh == 2 && !(p.pt == 0 && p.pl == 0)) return false;
    if (p.epi.res && p.epi.res == p.in && p.epi.res_Cs != p.Cs_in) return false;
    if (p.M % (p.Ho * p.Wo)) return false;
    if (p.Cs_in == 8) return p.sh == 1 && p.Ns <= 32;   // strip kernel only
    return p.Cs_in == 16 || p.Cs_in == 24 || p.Cs_in == 32 || p.Cs_in == 40 || p.Cs_in == 48;
}

bool launch_dwpw_thin(const ConvDev &p, cudaStream_t s) {
    if (!dwpw_thin_supported(p)) return false;
    if (launch_dwpw_strip(p, s)) return true;
    switch (p.Cs_in) {
        case 16: return launch_thin_cs<16>(p, s);
        case 24: return launch_thin_cs<24>(p, s);
        case 32: return launch_thin_cs<32>(p, s);
        case 40: return launch_thin_cs<40>(p, s);
        case 48: return launch_thin_cs<48>(p, s);
        default: return false;
    }
}


// Pointwise convs the thin kernel takes from the GEMM tile: 1x1 stride 1, Cs_in 16 or 32, N <= 32, maps >= 32x32,
// 4-aligned stores, residual (if any) vectorisable.
bool pw_thin_supported(const ConvDev &p) {
    static const bool disabled = getenv("ZB_NO_PW_THIN") && atoi(getenv("ZB_NO_PW_THIN")) != 0;
    if (disabled) return false;
    if (p.kh != 1 || p.kw != 1 || p.sh != 1 || p.sw != 1 || p.pt != 0 || p.pl != 0) return false;
    if (p.K != p.Cs_in || (p.Cs_in != 16 && p.Cs_in != 32) || p.Ns > 32 || p.Ns % 4 || p.Nstore % 4 || p.out_pix_stride % 4) return false;
    if (p.Ho * p.Wo < 1024 || p.M % (p.Ho * p.Wo)) return false;
    if (p.epi.res && (p.epi.res_Cs % 4)) return false;
    return true;
}

bool launch_pw_thin(const ConvDev &p, cudaStream_t s) {
    if (!pw_thin_supported(p)) return false;
    const int np = p.Ns <= 8 ? 8 : p.Ns <= 16 ? 16 : 32;
    if (p.Cs_in == 16) {
        if (np == 8) return launch_pw_thin_cfg<16, 8, 4>(p, s);
        if (np == 16) return launch_pw_thin_cfg<16, 16, 4>(p, s);
        return launch_pw_thin_cfg<16, 32, 2>(p, s);
    }
    if (np == 8) return launch_pw_thin_cfg<32, 8, 2>(p, s);
    if (np == 16) return launch_pw_thin_cfg<32, 16, 2>(p, s);
    return launch_pw_thin_cfg<32, 32, 2>(p, s);
}

// Heads with a 1x1 output map and little work (see dense_head_kernel); large N x M stays on the tcgen05 GEMM.
bool dense_head_supported(const ConvDev &p) {
    static const bool disabled = getenv("ZB_NO_DENSE_HEAD") && atoi(getenv("ZB_NO_DENSE_HEAD")) != 0;
    if (disabled) return false;
    if (p.Ho != 1 || p.Wo != 1 || p.kh != p.H || p.kw != p.W || p.pt != 0 || p.pl != 0) return false;
    if (p.K != p.kh * p.kw * p.Cs_in || p.K % 4 || p.in_img_stride % 4 || ((uintptr_t)p.in) % 16 || p.epi.res) return false;
    return (long long)p.M * p.K * p.Ns <= (64ll << 20);
}

bool launch_dense_head(const ConvDev &p, cudaStream_t s) {
    if (!dense_head_supported(p)) return false;
    g_launch_count++;
    constexpr int R = 4;
    ZB_KNAME("dense_head_kernel", R);
    launch_pdl(32, dense_head_kernel<R>, dim3((unsigned)((p.Ns + 31) / 32), (unsigned)((p.M + R - 1) / R)), dim3(256), 0, s, p);
    return true;
}

}  // namespace zb
