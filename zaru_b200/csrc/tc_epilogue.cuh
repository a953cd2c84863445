// Fused epilogue of the tcgen05 kernels (shared by kernels_tc.cu and kernels_tcb.cu): accumulator columns of one output
// pixel (one thread) -> + bias -> act1 -> + residual [channel-pad, 2x2 max-pool] -> act2 -> 128-bit stores.
#pragma once
#include "conv_common.cuh"
#include "tc_common.cuh"

namespace zb {
namespace {

using namespace tc;
constexpr int TCB_EPI_ROWS = 128;          // accumulator rows of a tile (UMMA M)

// Epilogue for `NC` consecutive accumulator columns of one output pixel (one thread).
// The residual of a column chunk is fetched FIRST (NC / 4 independent 128-bit loads in flight, issued before the
// TMEM read) instead of one dependent load per group of four: ncu had ~30 % of this kernel's stall samples on the
// `v += residual` FADDs waiting for their load.  Chunks of 16 columns keep v[] + rr[] at the register budget of
// the former 32-column chunk (3 CTAs per SM need <= 80 registers).
template <int NC>
struct ResidualPrefetch {
    float4 rr[NC / 4];
};
template <int NC>
__device__ __forceinline__ void tc_prefetch_residual(const ConvDev &p, int c0, int img, int oy, int ox, ResidualPrefetch<NC> &pre) {
    const EpiDev &e = p.epi;
    const bool vec = e.res && (e.res_Cs % 4) == 0;
#pragma unroll
    for (int h = 0; h < NC / 4; h++) {
        const int n = c0 + 4 * h;
        pre.rr[h] = (vec && n < p.Nstore) ? residual4_at(e, img, oy, ox, n) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
template <int NC>
__device__ __forceinline__ void tc_epilogue_cols(const ConvDev &p, const float (&acc)[NC], int c0, int img, int oy, int ox,
                                                 float *orow, bool vec_ok, const ResidualPrefetch<NC> &pre) {
    const EpiDev &e = p.epi;
#pragma unroll
    for (int h = 0; h < NC / 4; h++) {
        const int n = c0 + 4 * h;
        if (n >= p.Nstore) continue;
        float v[4] = {acc[4 * h], acc[4 * h + 1], acc[4 * h + 2], acc[4 * h + 3]};
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
                const float4 rr = pre.rr[h];
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

// ------------------------------------------------------------------------------------------------
// Tile epilogue THROUGH SHARED MEMORY.  The row-per-thread epilogue above walks its columns serially and a warp's 32
// stores of one instruction land in 32 different rows.  Here the accumulator tile goes TMEM -> registers -> shared memory
// (row-major, padded), then every thread takes (row, channel quad) items with the quad fastest: residual reads and output
// stores are coalesced 128-bit accesses and a thread's items are independent, so all its global loads are issued before
// the first one is consumed.  A thread keeps ONE channel quad for a whole column block, so bias and PReLU slopes are
// loaded once, and the row table holds ready-made element offsets - the first version spent ~200 instructions per item on
// index arithmetic and per-item guards (ncu: 90 % of the kernel's instructions were not FMAs).
//   s_rowinfo[m] = {element offset of the output pixel, element offset of its residual pixel}; out < 0: no pixel;
//   s_stage: 128 x TCE_STRIDE floats;  n0 = first output channel of this accumulator, NT = its column count (mult. of 16).
// Called by NTHREADS threads (tid = 0 .. NTHREADS - 1) that synchronise on barrier BAR_ID (0 = the whole CTA).
// ------------------------------------------------------------------------------------------------
constexpr int TCE_NB = 64;                 // columns per staging block
constexpr int TCE_STRIDE = TCE_NB + 4;     // padded row stride (floats): conflict-free 128-bit row-per-lane stores

struct TceRow {
    long long out_off, res_off;
};

template <int NTHREADS, int BAR_ID>
__device__ __forceinline__ void tce_sync() {
    if (BAR_ID == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"n"(BAR_ID), "n"(NTHREADS) : "memory");
}

// row table entry of output pixel (img, oy, ox); the residual offset addresses the pixel itself, or the top-left pixel of
// its 2x2 pooling window
__device__ __forceinline__ TceRow tce_row(const ConvDev &p, int img, int oy, int ox) {
    TceRow r;
    r.out_off = (long long)img * p.out_img_stride + ((long long)oy * p.Wo + ox) * p.out_pix_stride;
    const EpiDev &e = p.epi;
    r.res_off = 0;
    if (e.res) r.res_off = (long long)img * e.res_img_stride + (e.res_pool ? ((long long)(2 * oy) * e.res_W + 2 * ox) : ((long long)oy * e.res_W + ox)) * e.res_Cs;
    return r;
}

__device__ __forceinline__ void tce_act(float (&v)[4], int kind, float lo, float hi, const float4 &sl) {
    if (kind == ACT_NONE) return;
    if (kind == ACT_RELU) {
#pragma unroll
        for (int q = 0; q < 4; q++) v[q] = fmaxf(v[q], 0.0f);
    } else if (kind == ACT_PRELU) {
        v[0] = v[0] < 0.0f ? v[0] * sl.x : v[0];
        v[1] = v[1] < 0.0f ? v[1] * sl.y : v[1];
        v[2] = v[2] < 0.0f ? v[2] * sl.z : v[2];
        v[3] = v[3] < 0.0f ? v[3] * sl.w : v[3];
    } else if (kind == ACT_CLIP) {
#pragma unroll
        for (int q = 0; q < 4; q++) v[q] = fminf(fmaxf(v[q], lo), hi);
    } else {
#pragma unroll
        for (int q = 0; q < 4; q++) v[q] = 1.0f / (1.0f + expf(-v[q]));
    }
}

template <int NTHREADS = 256, int BAR_ID = 0>
__device__ __forceinline__ void tc_epilogue_tile(const ConvDev &p, uint32_t tmem, int n0, int NT, const TceRow *s_rowinfo,
                                                 float *s_stage, int tid) {
    const EpiDev &e = p.epi;
    const int warp = tid >> 5, lane = tid & 31;
    const int row_t = (warp & 3) * 32 + lane, grp = warp >> 2;      // grp: which 16-column chunks this warp moves
    const uint32_t tbase = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    // fast path: everything 128-bit (true for every block-to-block tensor; graph-output heads with odd widths take the
    // guarded path)
    const bool fast = (p.out_pix_stride % 4 == 0) && (p.Nstore % 4 == 0) && (p.out_img_stride % 4 == 0) && (p.Ns % 4 == 0) &&
                      ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) && (!e.res || (e.res_Cs % 4) == 0);
    const long long pool_dx = e.res_Cs, pool_dy = (long long)e.res_W * e.res_Cs;
    for (int cb = 0; cb < NT; cb += TCE_NB) {
        const int nb = min(TCE_NB, NT - cb);             // multiple of 16
        // 1. TMEM -> shared
        for (int j = grp; j < nb / 16; j += NTHREADS / 128) {
            float v[16];
            tmem_ld16(tbase + (uint32_t)(cb + 16 * j), v);
            float *dst = s_stage + row_t * TCE_STRIDE + 16 * j;
#pragma unroll
            for (int h = 0; h < 4; h++) *reinterpret_cast<float4 *>(dst + 4 * h) = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
        }
        tce_sync<NTHREADS, BAR_ID>();
        // 2. one channel quad per thread, rows r0, r0 + rstep, ...
        const int qb = nb >> 2;
        const int rstep = NTHREADS / qb;
        const int q = tid % qb, r0 = tid / qb;
        const int n = n0 + cb + 4 * q;
        if (r0 < rstep && n < p.Nstore) {
            if (fast) {
                const float4 bias = ldg4(e.bias + n);
                const float4 sl1 = e.act1.kind == ACT_PRELU ? ldg4(e.act1.slope + n) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 sl2 = e.act2.kind == ACT_PRELU ? ldg4(e.act2.slope + n) : make_float4(0.f, 0.f, 0.f, 0.f);
                const bool has_res = e.res && n < e.res_Cs;
                const float *resn = e.res + n;
                constexpr int MAXI = NTHREADS >= 512 ? 4 : 8;   // rows in flight per thread (512-thread CTAs run at 64 registers)
#pragma unroll 1
                for (int rb = r0; rb < TCB_EPI_ROWS; rb += MAXI * rstep) {
                    TceRow ri[MAXI];
                    float4 rr[MAXI];
#pragma unroll
                    for (int i = 0; i < MAXI; i++) {
                        const int r = rb + i * rstep;
                        ri[i].out_off = -1;
                        rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (r < TCB_EPI_ROWS) {
                            ri[i] = s_rowinfo[r];
                            if (ri[i].out_off >= 0 && has_res) {
                                const float *rp = resn + ri[i].res_off;
                                if (!e.res_pool) {
                                    rr[i] = ldg4(rp);
                                } else {
                                    const float4 a = ldg4(rp), b = ldg4(rp + pool_dx), c = ldg4(rp + pool_dy), d = ldg4(rp + pool_dy + pool_dx);
                                    rr[i] = make_float4(fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y)),
                                                        fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z)), fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w)));
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < MAXI; i++) {
                        if (ri[i].out_off < 0) continue;
                        const int r = rb + i * rstep;
                        const float4 a = *reinterpret_cast<const float4 *>(s_stage + r * TCE_STRIDE + 4 * q);
                        float v[4] = {a.x + bias.x, a.y + bias.y, a.z + bias.z, a.w + bias.w};
                        tce_act(v, e.act1.kind, e.act1.lo, e.act1.hi, sl1);
                        v[0] += rr[i].x, v[1] += rr[i].y, v[2] += rr[i].z, v[3] += rr[i].w;
                        tce_act(v, e.act2.kind, e.act2.lo, e.act2.hi, sl2);
                        *reinterpret_cast<float4 *>(p.out + ri[i].out_off + n) = make_float4(v[0], v[1], v[2], v[3]);
                    }
                }
            } else {
                // guarded path: scalar tails of bias / store, scalar residual
                for (int r = r0; r < TCB_EPI_ROWS; r += rstep) {
                    const TceRow ri = s_rowinfo[r];
                    if (ri.out_off < 0) continue;
                    const float4 a = *reinterpret_cast<const float4 *>(s_stage + r * TCE_STRIDE + 4 * q);
                    float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (n + k < p.Ns) v[k] += __ldg(e.bias + n + k);
                    act4(v, e.act1, n);
                    if (e.res) {
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            if (n + k >= e.res_Cs) continue;
                            const float *rp = e.res + ri.res_off + n + k;
                            v[k] += e.res_pool ? fmaxf(fmaxf(__ldg(rp), __ldg(rp + pool_dx)), fmaxf(__ldg(rp + pool_dy), __ldg(rp + pool_dy + pool_dx)))
                                               : __ldg(rp);
                        }
                    }
                    act4(v, e.act2, n);
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (n + k < p.Nstore) p.out[ri.out_off + n + k] = v[k];
                }
            }
        }
        tce_sync<NTHREADS, BAR_ID>();                   // the staging tile is overwritten by the next column block
    }
}

}  // namespace
}  // namespace zb
