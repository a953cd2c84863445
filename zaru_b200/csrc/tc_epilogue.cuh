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
// Tile epilogue THROUGH SHARED MEMORY.  The row-per-thread epilogue above walks its columns serially, every group of four
// waiting for its own residual / bias load (ncu on the first tile-block kernel: 45 % of all stall samples on those
// FADDs), and a warp's 32 stores of one instruction land in 32 different rows.  Here the accumulator tile goes
// TMEM -> registers -> shared memory (row-major, padded), then every thread takes (row, channel quad) items with the quad
// fastest: residual reads and output stores are coalesced 128-bit accesses, and a thread's items are independent, so
// all its global loads are issued before the first one is consumed.
//   s_rowinfo[m] = {img, oy, ox, ok} of accumulator row m;   s_stage: 128 x TCE_STRIDE floats;
//   n0 = first output channel of this CTA's accumulator, NT = its column count (multiple of 16).
// Must be called by all 256 threads of the CTA (contains __syncthreads).
// ------------------------------------------------------------------------------------------------
constexpr int TCE_NB = 64;                 // columns per staging block
constexpr int TCE_STRIDE = TCE_NB + 4;     // padded row stride (floats): conflict-free 128-bit row-per-lane stores

// NTHREADS = threads of the CTA (256 or 512); (row, quad) items per thread and block = 128 * 16 / NTHREADS
template <int NTHREADS = 256>
__device__ __forceinline__ void tc_epilogue_tile(const ConvDev &p, uint32_t tmem, int n0, int NT, const int4 *s_rowinfo,
                                                 float *s_stage, int tid) {
    constexpr int TCE_ITEMS = 128 * (TCE_NB / 4) / NTHREADS;
    const EpiDev &e = p.epi;
    const int warp = tid >> 5, lane = tid & 31;
    const int row_t = (warp & 3) * 32 + lane, half = warp >> 2;     // `half`: which 16-column chunks this warp moves (NTHREADS / 128 groups)
    const uint32_t tbase = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const bool vec_store = (p.out_pix_stride % 4 == 0) && (p.Nstore % 4 == 0) && (p.out_img_stride % 4 == 0) &&
                           ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    const bool vec_res = e.res && (e.res_Cs % 4) == 0;
    for (int cb = 0; cb < NT; cb += TCE_NB) {
        const int nb = min(TCE_NB, NT - cb);             // multiple of 16
        // 1. TMEM -> shared: 16-column chunks alternate between the two warp halves
        for (int j = half; j < nb / 16; j += NTHREADS / 128) {
            float v[16];
            tmem_ld16(tbase + (uint32_t)(cb + 16 * j), v);
            float *dst = s_stage + row_t * TCE_STRIDE + 16 * j;
#pragma unroll
            for (int h = 0; h < 4; h++) *reinterpret_cast<float4 *>(dst + 4 * h) = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
        }
        __syncthreads();
        // 2. (row, quad) items, quad fastest
        const int qb = nb >> 2, total = TCB_EPI_ROWS * qb;
        int rows[TCE_ITEMS], ns[TCE_ITEMS];
        int4 ri[TCE_ITEMS];
        float4 rr[TCE_ITEMS];
#pragma unroll
        for (int i = 0; i < TCE_ITEMS; i++) {            // all residual loads first
            const int it = tid + i * NTHREADS;
            rows[i] = -1;
            rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (it < total) {
                const int r = it / qb, q = it - r * qb;
                const int n = n0 + cb + 4 * q;
                ri[i] = s_rowinfo[r];
                if (ri[i].w && n < p.Nstore) {
                    rows[i] = r, ns[i] = n;
                    if (vec_res) rr[i] = residual4_at(e, ri[i].x, ri[i].y, ri[i].z, n);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < TCE_ITEMS; i++) {
            if (rows[i] < 0) continue;
            const int n = ns[i];
            const float4 a = *reinterpret_cast<const float4 *>(s_stage + rows[i] * TCE_STRIDE + (n - n0 - cb));
            float v[4] = {a.x, a.y, a.z, a.w};
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
                if (vec_res) {
                    v[0] += rr[i].x, v[1] += rr[i].y, v[2] += rr[i].z, v[3] += rr[i].w;
                } else {
#pragma unroll
                    for (int q = 0; q < 4; q++) v[q] += residual_at(e, ri[i].x, ri[i].y, ri[i].z, n + q);
                }
            }
            act4(v, e.act2, n);
            float *orow = p.out + (long long)ri[i].x * p.out_img_stride + ((long long)ri[i].y * p.Wo + ri[i].z) * p.out_pix_stride;
            if (vec_store) {
                *reinterpret_cast<float4 *>(orow + n) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (n + q < p.Nstore) orow[n + q] = v[q];
            }
        }
        __syncthreads();                                 // the staging tile is overwritten by the next column block
    }
}

}  // namespace
}  // namespace zb
