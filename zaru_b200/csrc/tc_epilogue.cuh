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
    long long out_off, res_off;            // BYTE offsets of the output pixel / its residual pixel; out_off < 0: no pixel
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
    r.out_off = 4 * ((long long)img * p.out_img_stride + ((long long)oy * p.Wo + ox) * p.out_pix_stride);
    const EpiDev &e = p.epi;
    r.res_off = 0;
    if (e.res)
        r.res_off = 4 * ((long long)img * e.res_img_stride +
                         (e.res_pool ? ((long long)(2 * oy) * e.res_W + 2 * ox) : ((long long)oy * e.res_W + ox)) * e.res_Cs);
    return r;
}

// activation of one channel quad.  The kinds arrive as flags the caller derived ONCE (an if-chain on the raw kind was
// turned into a jump table - LDC + BRX per call - by the compiler)
struct TceAct {
    bool relu, prelu, other;
    int kind;
    float lo, hi;
};
__device__ __forceinline__ TceAct tce_act_of(const ActDev &a) {
    TceAct t;
    t.relu = a.kind == ACT_RELU, t.prelu = a.kind == ACT_PRELU;
    t.other = a.kind != ACT_NONE && !t.relu && !t.prelu;
    t.kind = a.kind, t.lo = a.lo, t.hi = a.hi;
    return t;
}
__device__ __forceinline__ void tce_act(float4 &v, const TceAct &a, const float4 &sl) {
    if (a.prelu) {
        v.x = v.x < 0.0f ? v.x * sl.x : v.x;
        v.y = v.y < 0.0f ? v.y * sl.y : v.y;
        v.z = v.z < 0.0f ? v.z * sl.z : v.z;
        v.w = v.w < 0.0f ? v.w * sl.w : v.w;
    } else if (a.relu) {
        v.x = fmaxf(v.x, 0.0f), v.y = fmaxf(v.y, 0.0f), v.z = fmaxf(v.z, 0.0f), v.w = fmaxf(v.w, 0.0f);
    } else if (a.other) {
        if (a.kind == ACT_CLIP) {
            v.x = fminf(fmaxf(v.x, a.lo), a.hi), v.y = fminf(fmaxf(v.y, a.lo), a.hi);
            v.z = fminf(fmaxf(v.z, a.lo), a.hi), v.w = fminf(fmaxf(v.w, a.lo), a.hi);
        } else {
            v.x = 1.0f / (1.0f + expf(-v.x)), v.y = 1.0f / (1.0f + expf(-v.y));
            v.z = 1.0f / (1.0f + expf(-v.z)), v.w = 1.0f / (1.0f + expf(-v.w));
        }
    }
}

// Fast path, split in two so that the global loads fly while the accumulator is still on its way:
//   tce_fetch  - row table entries + residual loads of NI rows r0, r0 + rstep, ... of one channel quad (issued, not consumed)
//   tce_finish - staged accumulator + bias -> act1 -> + residual -> act2 -> 128-bit store
// Rows without a pixel (out_off < 0) run the arithmetic on zeros and skip the store: no divergent control flow in the loop.
template <int NI>
struct TceFetch {
    float4 rr[NI];                 // residual quads in flight (the row offsets are re-read from the table: registers)
};
template <int NI, bool GUARD>
__device__ __forceinline__ void tce_fetch(const EpiDev &e, const TceRow *rt, int r0, int rstep, const char *resb, bool has_res, TceFetch<NI> &f,
                                          int limit = TCB_EPI_ROWS) {
#pragma unroll
    for (int i = 0; i < NI; i++) f.rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!has_res) return;
    if (!e.res_pool) {
#pragma unroll
        for (int i = 0; i < NI; i++) {
            if (GUARD && r0 + i * rstep >= limit) continue;
            const TceRow t = rt[i * rstep];
            if (t.out_off >= 0) f.rr[i] = ldg4(reinterpret_cast<const float *>(resb + t.res_off));
        }
    } else {                                           // 2x2 max-pool of the residual (stride-2 blocks)
        const long long pool_dx = 4ll * e.res_Cs, pool_dy = 4ll * e.res_W * e.res_Cs;
#pragma unroll
        for (int i = 0; i < NI; i++) {
            if (GUARD && r0 + i * rstep >= limit) continue;
            const TceRow t = rt[i * rstep];
            if (t.out_off < 0) continue;
            const char *rp = resb + t.res_off;
            const float4 a = ldg4(reinterpret_cast<const float *>(rp)), b = ldg4(reinterpret_cast<const float *>(rp + pool_dx));
            const float4 c = ldg4(reinterpret_cast<const float *>(rp + pool_dy)), d = ldg4(reinterpret_cast<const float *>(rp + pool_dy + pool_dx));
            f.rr[i] = make_float4(fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y)),
                                  fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z)), fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w)));
        }
    }
}
template <int NI, bool GUARD>
__device__ __forceinline__ void tce_finish(const TceFetch<NI> &f, const TceRow *rt, const float *stg, int r0, int rstep, char *outb,
                                           const float4 &bias, const TceAct &a1, const TceAct &a2, const float4 &sl1, const float4 &sl2,
                                           int limit = TCB_EPI_ROWS) {
#pragma unroll
    for (int i = 0; i < NI; i++) {
        if (GUARD && r0 + i * rstep >= limit) continue;
        const float4 a = *reinterpret_cast<const float4 *>(stg + i * rstep * TCE_STRIDE);
        const long long oo = rt[i * rstep].out_off;
        float4 v = make_float4(a.x + bias.x, a.y + bias.y, a.z + bias.z, a.w + bias.w);
        tce_act(v, a1, sl1);
        v.x += f.rr[i].x, v.y += f.rr[i].y, v.z += f.rr[i].z, v.w += f.rr[i].w;
        tce_act(v, a2, sl2);
        if (oo >= 0) *reinterpret_cast<float4 *>(outb + oo) = v;
    }
}

struct TceNoWait {
    __device__ __forceinline__ void operator()() const {}
};

// `ready()` is called once, after the first column block's global loads have been issued and before the accumulator is
// read: callers park their wait for the tile's last MMAs there.
template <int NTHREADS = 256, int BAR_ID = 0, class Ready = TceNoWait>
__device__ __forceinline__ void tc_epilogue_tile(const ConvDev &p, uint32_t tmem, int n0, int NT, const TceRow *s_rowinfo,
                                                 float *s_stage, int tid, Ready ready = Ready()) {
    const EpiDev &e = p.epi;
    const int warp = tid >> 5, lane = tid & 31;
    const int row_t = (warp & 3) * 32 + lane, grp = warp >> 2;      // grp: which 16-column chunks this warp moves
    const uint32_t tbase = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    // fast path: everything 128-bit (true for every block-to-block tensor; graph-output heads with odd widths take the
    // guarded path)
    const bool fast = (p.out_pix_stride % 4 == 0) && (p.Nstore % 4 == 0) && (p.out_img_stride % 4 == 0) && (p.Ns % 4 == 0) &&
                      ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) && (!e.res || (e.res_Cs % 4) == 0);
    const TceAct a1 = tce_act_of(e.act1), a2 = tce_act_of(e.act2);
    // rows in flight per thread: one batch covers the tile when a thread has <= 8 rows (512-thread CTAs run at 64 registers: 4)
    constexpr int NI = NTHREADS >= 512 ? 4 : 8;
    for (int cb = 0; cb < NT; cb += TCE_NB) {
        const int nb = min(TCE_NB, NT - cb);             // multiple of 16
        // one channel quad per thread, rows r0, r0 + rstep, ...
        const int qb = nb >> 2;                          // 4, 8, 12 or 16 quads
        int q, r0, rstep;
        if (qb == 16) q = tid & 15, r0 = tid >> 4, rstep = NTHREADS / 16;
        else if (qb == 8) q = tid & 7, r0 = tid >> 3, rstep = NTHREADS / 8;
        else if (qb == 4) q = tid & 3, r0 = tid >> 2, rstep = NTHREADS / 4;
        else q = tid % 12, r0 = tid / 12, rstep = NTHREADS / 12;
        const int n = n0 + cb + 4 * q;
        const bool active = r0 < rstep && n < p.Nstore;
        // 0. (fast path) everything that comes from global memory for the first batch of rows
        TceFetch<NI> f;
        float4 bias = make_float4(0.f, 0.f, 0.f, 0.f), sl1 = bias, sl2 = bias;
        const bool has_res = e.res && n < e.res_Cs;
        const char *resb = reinterpret_cast<const char *>(e.res + n);
        char *outb = reinterpret_cast<char *>(p.out + n);
        if (fast && active) {
            bias = ldg4(e.bias + n);
            if (a1.prelu) sl1 = ldg4(e.act1.slope + n);
            if (a2.prelu) sl2 = ldg4(e.act2.slope + n);
            tce_fetch<NI, true>(e, s_rowinfo + r0, r0, rstep, resb, has_res, f);
        }
        if (cb == 0) ready();
        // 1. TMEM -> shared
        for (int j = grp; j < nb / 16; j += NTHREADS / 128) {
            float v[16];
            tmem_ld16(tbase + (uint32_t)(cb + 16 * j), v);
            float *dst = s_stage + row_t * TCE_STRIDE + 16 * j;
#pragma unroll
            for (int h = 0; h < 4; h++) *reinterpret_cast<float4 *>(dst + 4 * h) = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
        }
        tce_sync<NTHREADS, BAR_ID>();
        // 2. finish
        if (active) {
            if (fast) {
                const float *stg = s_stage + r0 * TCE_STRIDE + 4 * q;
                tce_finish<NI, true>(f, s_rowinfo + r0, stg, r0, rstep, outb, bias, a1, a2, sl1, sl2);
#pragma unroll 1
                for (int rb = r0 + NI * rstep; rb < TCB_EPI_ROWS; rb += NI * rstep) {     // further batches (small CTAs / few quads)
                    tce_fetch<NI, true>(e, s_rowinfo + rb, rb, rstep, resb, has_res, f);
                    tce_finish<NI, true>(f, s_rowinfo + rb, s_stage + rb * TCE_STRIDE + 4 * q, rb, rstep, outb, bias, a1, a2, sl1, sl2);
                }
            } else {
                // guarded path: scalar tails of bias / store, scalar residual
                const long long pool_dx = e.res_Cs, pool_dy = (long long)e.res_W * e.res_Cs;
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
                            const float *rp = e.res + (ri.res_off >> 2) + n + k;
                            v[k] += e.res_pool ? fmaxf(fmaxf(__ldg(rp), __ldg(rp + pool_dx)), fmaxf(__ldg(rp + pool_dy), __ldg(rp + pool_dy + pool_dx)))
                                               : __ldg(rp);
                        }
                    }
                    act4(v, e.act2, n);
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (n + k < p.Nstore) p.out[(ri.out_off >> 2) + n + k] = v[k];
                }
            }
        }
        tce_sync<NTHREADS, BAR_ID>();                   // the staging tile is overwritten by the next column block
    }
}

// ------------------------------------------------------------------------------------------------
// The same epilogue with QUARTER-LOCAL barriers.  A warp can only read its own 32-lane quarter of the accumulator, so rows
// 32 q .. 32 q + 31 of the staging tile are written by the NTHREADS / 128 warps with (warp & 3) == q - and if exactly those
// warps also finish those rows, nobody needs the other quarters: the CTA-wide barrier between staging and finishing becomes
// four independent barriers of NTHREADS / 4 threads (ids 1 .. 4), and a warp that arrives late (the one that issued the MMAs,
// the one whose residual load was slow) holds up its quarter instead of the CTA.  ncu on the 5x5 palm blocks: 17 % of the warp
// samples sat on the two CTA-wide epilogue barriers.  There is NO barrier at the end: the caller orders the reuse of the staging
// tile (it aliases the A tile) and of the row table with its own CTA barrier before it overwrites them.
// Fast path only (everything 128-bit); returns false without touching anything when the tensor needs the guarded path.
// ------------------------------------------------------------------------------------------------
template <int NTHREADS>
__device__ __forceinline__ void tce_quarter_sync(int q4) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + q4), "n"(NTHREADS / 4) : "memory");
}

__device__ __forceinline__ bool tce_fast_ok(const ConvDev &p) {
    const EpiDev &e = p.epi;
    return (p.out_pix_stride % 4 == 0) && (p.Nstore % 4 == 0) && (p.out_img_stride % 4 == 0) && (p.Ns % 4 == 0) &&
           ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) && (!e.res || (e.res_Cs % 4) == 0);
}

template <int NTHREADS, class Ready>
__device__ __forceinline__ void tc_epilogue_tile_quarters(const ConvDev &p, uint32_t tmem, int n0, int NT, const TceRow *s_rowinfo,
                                                          float *s_stage, int tid, Ready ready) {
    const EpiDev &e = p.epi;
    constexpr int GT = NTHREADS / 4;                    // threads per quarter group
    const int warp = tid >> 5, lane = tid & 31;
    const int q4 = warp & 3, grp = warp >> 2;           // lane quarter; which 16-column chunks this warp moves
    const int row_t = q4 * 32 + lane;
    const uint32_t tbase = tmem + ((uint32_t)(q4 * 32) << 16);
    const TceAct a1 = tce_act_of(e.act1), a2 = tce_act_of(e.act2);
    constexpr int NI = NTHREADS >= 512 ? 4 : 8;
    const int t = grp * 32 + lane;                      // index inside the quarter group
    const TceRow *rt_q = s_rowinfo + q4 * 32;
    float *stage_q = s_stage + q4 * 32 * TCE_STRIDE;
    for (int cb = 0; cb < NT; cb += TCE_NB) {
        const int nb = min(TCE_NB, NT - cb);            // multiple of 16
        const int qb = nb >> 2;                         // 4, 8, 12 or 16 quads per row
        int q, rl, rstep;                               // this thread's quad, first local row, local row step
        if (qb == 16) q = t & 15, rl = t >> 4, rstep = GT / 16;
        else if (qb == 8) q = t & 7, rl = t >> 3, rstep = GT / 8;
        else if (qb == 4) q = t & 3, rl = t >> 2, rstep = GT / 4;
        else q = t % 12, rl = t / 12, rstep = GT / 12;
        const int n = n0 + cb + 4 * q;
        const bool active = rl < rstep && n < p.Nstore;
        TceFetch<NI> f;
        float4 bias = make_float4(0.f, 0.f, 0.f, 0.f), sl1 = bias, sl2 = bias;
        const bool has_res = e.res && n < e.res_Cs;
        const char *resb = reinterpret_cast<const char *>(e.res + n);
        char *outb = reinterpret_cast<char *>(p.out + n);
        if (active) {
            bias = ldg4(e.bias + n);
            if (a1.prelu) sl1 = ldg4(e.act1.slope + n);
            if (a2.prelu) sl2 = ldg4(e.act2.slope + n);
            tce_fetch<NI, true>(e, rt_q + rl, rl, rstep, resb, has_res, f, 32);
        }
        if (cb == 0) ready();
        else tce_quarter_sync<NTHREADS>(q4);            // the previous column block's rows of this quarter have been finished
        for (int j = grp; j < nb / 16; j += NTHREADS / 128) {
            float v[16];
            tmem_ld16(tbase + (uint32_t)(cb + 16 * j), v);
            float *dst = s_stage + row_t * TCE_STRIDE + 16 * j;
#pragma unroll
            for (int h = 0; h < 4; h++) *reinterpret_cast<float4 *>(dst + 4 * h) = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
        }
        tce_quarter_sync<NTHREADS>(q4);
        if (active) {
            tce_finish<NI, true>(f, rt_q + rl, stage_q + rl * TCE_STRIDE + 4 * q, rl, rstep, outb, bias, a1, a2, sl1, sl2, 32);
#pragma unroll 1
            for (int rb = rl + NI * rstep; rb < 32; rb += NI * rstep) {
                tce_fetch<NI, true>(e, rt_q + rb, rb, rstep, resb, has_res, f, 32);
                tce_finish<NI, true>(f, rt_q + rb, stage_q + rb * TCE_STRIDE + 4 * q, rb, rstep, outb, bias, a1, a2, sl1, sl2, 32);
            }
        }
    }
}

}  // namespace
}  // namespace zb
