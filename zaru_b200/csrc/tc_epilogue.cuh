// Fused epilogue of the tcgen05 kernels (shared by kernels_tc.cu and kernels_tcb.cu): accumulator columns of one output
// pixel (one thread) -> + bias -> act1 -> + residual [channel-pad, 2x2 max-pool] -> act2 -> 128-bit stores.
#pragma once
#include "conv_common.cuh"

namespace zb {
namespace {

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

}  // namespace
}  // namespace zb
