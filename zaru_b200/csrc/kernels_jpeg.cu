// JPEG back end on the device: dequantisation + inverse DCT, chroma upsampling and YCbCr -> RGB, straight into the RGBA8
// frame pool.  Replaces the pixel half of `zaru_image::jpeg::decode_jpeg` (crates/zaru-image/src/jpeg.rs:107-222).
//
// The reference can decode with five different libraries (ZARU_JPEG_BACKEND) whose pixels differ in the last bits; this
// implementation is BIT-EXACT with ONE of them - libjpeg-turbo's default pipeline, the one `turbojpeg` / `mozjpeg` use and
// the one OpenCV and Pillow ship (the oracle of tests/test_jpeg.py):
//   * "islow" integer inverse DCT (jidctint.c: 13-bit constants, two passes, descale with rounding),
//   * "fancy" (triangle-filter) chroma upsampling: h2v1 for 4:2:2, h2v2 for 4:2:0 (jdsample.c), edge rows / columns
//     replicated the way the decompressor's context rows do,
//   * YCbCr -> RGB with 16-bit fixed-point tables (jdcolor.c: 1.402, 0.34414, 0.71414, 1.772).
// Everything is integer arithmetic, so "bit-exact" is literal.
#include <cuda_runtime.h>

#include "jpeg_host.h"
#include "kernels.h"

namespace zb {
namespace {

constexpr int CONST_BITS = 13, PASS1_BITS = 2;
constexpr int FIX_0_298631336 = 2446, FIX_0_390180644 = 3196, FIX_0_541196100 = 4433, FIX_0_765366865 = 6270, FIX_0_899976223 = 7373,
              FIX_1_175875602 = 9633, FIX_1_501321110 = 12299, FIX_1_847759065 = 15137, FIX_1_961570560 = 16069, FIX_2_053119869 = 16819,
              FIX_2_562915447 = 20995, FIX_3_072711026 = 25172;

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// one 1-D pass of jidctint.c on 8 dequantised values; SHIFT = descale amount of this pass
template <int SHIFT, bool FIRST>
__device__ __forceinline__ void idct_1d(const int (&in)[8], int (&out)[8]) {
    int z1, z2, z3, z4, z5, tmp0, tmp1, tmp2, tmp3, tmp10, tmp11, tmp12, tmp13;
    z2 = in[2], z3 = in[6];
    z1 = (z2 + z3) * FIX_0_541196100;
    tmp2 = z1 + z3 * (-FIX_1_847759065);
    tmp3 = z1 + z2 * FIX_0_765366865;
    z2 = in[0], z3 = in[4];
    tmp0 = (z2 + z3) << CONST_BITS;
    tmp1 = (z2 - z3) << CONST_BITS;
    tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = in[7], tmp1 = in[5], tmp2 = in[3], tmp3 = in[1];
    z1 = tmp0 + tmp3, z2 = tmp1 + tmp2, z3 = tmp0 + tmp2, z4 = tmp1 + tmp3;
    z5 = (z3 + z4) * FIX_1_175875602;
    tmp0 *= FIX_0_298631336, tmp1 *= FIX_2_053119869, tmp2 *= FIX_3_072711026, tmp3 *= FIX_1_501321110;
    z1 *= -FIX_0_899976223, z2 *= -FIX_2_562915447, z3 *= -FIX_1_961570560, z4 *= -FIX_0_390180644;
    z3 += z5, z4 += z5;
    tmp0 += z1 + z3, tmp1 += z2 + z4, tmp2 += z2 + z3, tmp3 += z1 + z4;
    out[0] = descale(tmp10 + tmp3, SHIFT), out[7] = descale(tmp10 - tmp3, SHIFT);
    out[1] = descale(tmp11 + tmp2, SHIFT), out[6] = descale(tmp11 - tmp2, SHIFT);
    out[2] = descale(tmp12 + tmp1, SHIFT), out[5] = descale(tmp12 - tmp1, SHIFT);
    out[3] = descale(tmp13 + tmp0, SHIFT), out[4] = descale(tmp13 - tmp0, SHIFT);
}

struct JpegPlanes {
    const uint32_t *start;      // per flat block: first byte of its non-zero coefficients in `stream`
    const uint8_t *count;       // ... and how many there are
    const uint8_t *stream;      // {natural-order index, value lo, value hi} triples
    uint8_t *plane[3];          // decoded samples per component, [blocks_h * 8][blocks_w * 8]
    int blocks_w[3], blocks_h[3];
    int block_base[4];          // prefix sums of blocks per component (flat block index -> component)
    unsigned short qt[3][64];
    int ncomp;
};

// One thread per 8x8 block: pass 1 over columns (workspace in shared memory), pass 2 over rows, range limit.
__global__ void __launch_bounds__(64) jpeg_idct_kernel(const JpegPlanes P) {
    __shared__ int ws[64][65];
    const int b = blockIdx.x * 64 + threadIdx.x;
    if (b >= P.block_base[P.ncomp]) return;
    int c = 0;
    while (c + 1 < P.ncomp && b >= P.block_base[c + 1]) c++;
    const int bi = b - P.block_base[c];
    const int by = bi / P.blocks_w[c], bx = bi - by * P.blocks_w[c];
    int *w = ws[threadIdx.x];
    // expand the sparse coefficients (dequantised on the way) into the workspace
    for (int k = 0; k < 64; k++) w[k] = 0;
    {
        const uint8_t *sp = P.stream + P.start[b];
        const int n = P.count[b];
        for (int k = 0; k < n; k++) {
            const int idx = sp[3 * k] & 63;
            const int val = (int)(short)((unsigned short)sp[3 * k + 1] | ((unsigned short)sp[3 * k + 2] << 8));
            w[idx] = val * (int)P.qt[c][idx];
        }
    }
    for (int col = 0; col < 8; col++) {
        int v[8], o[8];
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = w[r * 8 + col];
        idct_1d<CONST_BITS - PASS1_BITS, true>(v, o);
#pragma unroll
        for (int r = 0; r < 8; r++) w[r * 8 + col] = o[r];
    }
    uint8_t *out = P.plane[c] + ((size_t)by * 8) * (P.blocks_w[c] * 8) + bx * 8;
    for (int r = 0; r < 8; r++) {
        int v[8], o[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = w[r * 8 + k];
        idct_1d<CONST_BITS + PASS1_BITS + 3, false>(v, o);
        unsigned lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const unsigned s = (unsigned)min(max(o[k] + 128, 0), 255);      // range_limit[]
            if (k < 4) lo |= s << (8 * k);
            else hi |= s << (8 * (k - 4));
        }
        *reinterpret_cast<uint2 *>(out + (size_t)r * (P.blocks_w[c] * 8)) = make_uint2(lo, hi);
    }
}

struct JpegOut {
    const uint8_t *plane[3];
    int pw[3];                  // plane row pitch (bytes)
    int width, height;          // image size
    int cw, ch;                 // chroma "downsampled" size: ceil(width / hmax), ceil(height / vmax)
    int hs, vs;                 // luma sampling factors (1 or 2)
    int ncomp;
    uint8_t *rgba;              // destination frame
    long long row_stride;
};

// chroma sample at full resolution (x, y): jdsample.c fancy upsampling
__device__ __forceinline__ int chroma_at(const uint8_t *pl, int pitch, int cw, int ch, int hs, int vs, int x, int y) {
    if (hs == 1 && vs == 1) return pl[(size_t)y * pitch + x];
    const int cx = x >> 1;
    if (cw <= 2) return pl[(size_t)(vs == 2 ? (y >> 1) : y) * pitch + cx];   // jdsample.c: fancy upsampling needs downsampled_width > 2
    if (vs == 1) {                                   // h2v1_fancy_upsample
        const uint8_t *row = pl + (size_t)y * pitch;
        const int cur = row[cx];
        if (cw == 1) return cur;
        if (x & 1) return cx + 1 < cw ? (cur * 3 + row[cx + 1] + 2) >> 2 : cur;
        return cx > 0 ? (cur * 3 + row[cx - 1] + 1) >> 2 : cur;
    }
    // h2v2_fancy_upsample: nearer chroma row weighted 3, the other 1; rows replicated at the top / bottom edge
    const int cy = y >> 1;
    const int oy = (y & 1) ? min(cy + 1, ch - 1) : max(cy - 1, 0);
    const uint8_t *r0 = pl + (size_t)cy * pitch, *r1 = pl + (size_t)oy * pitch;
    const int thiscol = r0[cx] * 3 + r1[cx];
    if (x & 1) {
        if (cx + 1 < cw) return (thiscol * 3 + (r0[cx + 1] * 3 + r1[cx + 1]) + 7) >> 4;
        return (thiscol * 4 + 7) >> 4;
    }
    if (cx > 0) return (thiscol * 3 + (r0[cx - 1] * 3 + r1[cx - 1]) + 8) >> 4;
    return (thiscol * 4 + 8) >> 4;
}

__global__ void __launch_bounds__(256) jpeg_color_kernel(const JpegOut o) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= o.width || y >= o.height) return;
    const int Y = o.plane[0][(size_t)y * o.pw[0] + x];
    int r = Y, g = Y, b = Y;
    if (o.ncomp == 3) {
        const int cb = chroma_at(o.plane[1], o.pw[1], o.cw, o.ch, o.hs, o.vs, x, y) - 128;
        const int cr = chroma_at(o.plane[2], o.pw[2], o.cw, o.ch, o.hs, o.vs, x, y) - 128;
        // jdcolor.c build_ycc_rgb_table: SCALEBITS 16, ONE_HALF 32768
        const int cr_r = (91881 * cr + 32768) >> 16, cb_b = (116130 * cb + 32768) >> 16;
        const int cr_g = -46802 * cr, cb_g = -22554 * cb + 32768;
        r = min(max(Y + cr_r, 0), 255);
        g = min(max(Y + ((cb_g + cr_g) >> 16), 0), 255);
        b = min(max(Y + cb_b, 0), 255);
    }
    *reinterpret_cast<unsigned *>(o.rgba + (size_t)y * o.row_stride + 4ll * x) = (unsigned)r | ((unsigned)g << 8) | ((unsigned)b << 16) | 0xff000000u;
}

}  // namespace

size_t jpeg_plane_bytes(const JpegHeader &h) {
    size_t n = 0;
    for (int c = 0; c < h.ncomp; c++) n += (size_t)h.blocks_w[c] * h.blocks_h[c] * 64;
    return n;
}

// start_dev / count_dev / stream_dev: one image's sparse coefficients on the device; plane_dev: scratch for its decoded
// component planes (jpeg_plane_bytes); rgba: destination frame.
void launch_jpeg_decode(const JpegHeader &h, const uint32_t *start_dev, const uint8_t *count_dev, const uint8_t *stream_dev,
                        uint8_t *plane_dev, uint8_t *rgba, long long row_stride, cudaStream_t s) {
    JpegPlanes P{};
    JpegOut O{};
    P.ncomp = O.ncomp = h.ncomp;
    P.start = start_dev, P.count = count_dev, P.stream = stream_dev;
    size_t po = 0;
    int blocks = 0;
    for (int c = 0; c < h.ncomp; c++) {
        P.plane[c] = plane_dev + po;
        O.plane[c] = plane_dev + po;
        P.blocks_w[c] = h.blocks_w[c], P.blocks_h[c] = h.blocks_h[c];
        O.pw[c] = h.blocks_w[c] * 8;
        P.block_base[c] = blocks;
        for (int k = 0; k < 64; k++) P.qt[c][k] = h.qt[h.tq[c]][k];
        const size_t nb = (size_t)h.blocks_w[c] * h.blocks_h[c];
        po += nb * 64, blocks += (int)nb;
    }
    for (int c = h.ncomp; c < 4; c++) P.block_base[c] = blocks;
    O.width = h.width, O.height = h.height;
    O.hs = h.hmax, O.vs = h.vmax;
    O.cw = (h.width + h.hmax - 1) / h.hmax, O.ch = (h.height + h.vmax - 1) / h.vmax;
    O.rgba = rgba, O.row_stride = row_stride;
    g_launch_count += 2;
    ZB_KNAME("jpeg_idct_kernel");
    jpeg_idct_kernel<<<(blocks + 63) / 64, 64, 0, s>>>(P);
    ZB_KNAME("jpeg_color_kernel");
    jpeg_color_kernel<<<dim3((h.width + 31) / 32, (h.height + 7) / 8), 256, 0, s>>>(O);
}

}  // namespace zb
