// Baseline JPEG front end (host): marker parsing + Huffman entropy decoding into quantised DCT coefficient blocks.
// What replaces `zaru_image::jpeg::decode_jpeg` (crates/zaru-image/src/jpeg.rs:107-222: zune-jpeg / mozjpeg / turbojpeg /
// jpeg-decoder behind ZARU_JPEG_BACKEND) for MJPG ingest (crates/zaru/src/video/webcam.rs:287, httpcam.rs:76): the
// sequential, data-dependent bit parsing stays on the host, everything per-pixel (dequantisation, inverse DCT, chroma
// upsampling, YCbCr -> RGB) runs on the device (kernels_jpeg.cu) straight into the RGBA8 frame pool.
//
// Scope: baseline sequential DCT (SOF0), 8-bit, Huffman, 1 or 3 components, luma sampling 1x1 / 2x1 / 2x2 with 1x1 chroma
// (4:4:4, 4:2:2, 4:2:0 - what webcams send), restart intervals.  Progressive (SOF2), arithmetic coding, 12-bit and CMYK
// are rejected with a message ("unsupported op ..." -> ZB_ERR_UNSUPPORTED_OP).
#include "jpeg_host.h"

#include <cstring>
#include <stdexcept>

namespace zb {
namespace {

[[noreturn]] void bad(const std::string &what) { throw std::runtime_error("jpeg: " + what); }
[[noreturn]] void unsupported(const std::string &what) { throw std::runtime_error("unsupported op: JPEG " + what); }

const uint8_t ZIGZAG[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                            41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                            30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {
    bool present = false;
    // canonical decoding (ITU T.81 F.2.2.3): per code length the smallest code, the largest code and the value index
    int mincode[17], maxcode[18], valptr[17];
    uint8_t vals[256];
    // 9-bit lookahead: (length << 8) | value, 0 = longer code
    uint16_t look[512];
    void build(const uint8_t counts[16], const uint8_t *symbols, int nsym) {
        memcpy(vals, symbols, nsym);
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            valptr[l] = k;
            mincode[l] = code;
            code += counts[l - 1];
            k += counts[l - 1];
            maxcode[l] = counts[l - 1] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        memset(look, 0, sizeof look);
        code = 0, k = 0;
        for (int l = 1; l <= 9; l++) {
            for (int i = 0; i < counts[l - 1]; i++, k++, code++) {
                const int first = code << (9 - l);
                for (int f = 0; f < (1 << (9 - l)); f++) look[first + f] = (uint16_t)((l << 8) | symbols[k]);
            }
            code <<= 1;
        }
        present = true;
    }
};

struct BitReader {
    const uint8_t *p, *end;
    uint64_t acc = 0;
    int bits = 0;
    bool hit_marker = false;
    BitReader(const uint8_t *b, const uint8_t *e) : p(b), end(e) {}
    void fill() {
        while (bits <= 56) {
            int byte = 0;
            if (!hit_marker && p < end) {
                byte = *p;
                if (byte == 0xFF) {
                    if (p + 1 < end && p[1] == 0x00) {
                        p += 2;
                    } else {
                        hit_marker = true;     // a marker ends the entropy-coded segment: feed zeros from here on
                        byte = 0;
                    }
                } else {
                    p++;
                }
            }
            acc |= (uint64_t)byte << (56 - bits);
            bits += 8;
        }
    }
    int peek(int n) {
        if (bits < n) fill();
        return (int)(acc >> (64 - n));
    }
    void skip(int n) {
        acc <<= n;
        bits -= n;
    }
    int get(int n) {
        if (n == 0) return 0;
        const int v = peek(n);
        skip(n);
        return v;
    }
    void reset_at(const uint8_t *np) {
        p = np;
        acc = 0;
        bits = 0;
        hit_marker = false;
    }
};

inline int decode_symbol(BitReader &br, const HuffTable &t) {
    const int look = t.look[br.peek(9)];
    if (look) {
        br.skip(look >> 8);
        return look & 255;
    }
    int code = br.peek(16);
    for (int l = 10; l <= 16; l++) {
        const int c = code >> (16 - l);
        if (t.maxcode[l] >= 0 && c <= t.maxcode[l] && c >= t.mincode[l]) {
            br.skip(l);
            return t.vals[t.valptr[l] + c - t.mincode[l]];
        }
    }
    bad("invalid Huffman code");
}

inline int extend(int v, int n) { return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v; }   // T.81 F.2.2.1 EXTEND

uint16_t be16(const uint8_t *p) { return (uint16_t)((p[0] << 8) | p[1]); }

}  // namespace

JpegHeader jpeg_parse_header(const uint8_t *data, size_t len) {
    if (len < 4 || data[0] != 0xFF || data[1] != 0xD8) bad("missing SOI marker");
    JpegHeader h;
    size_t i = 2;
    while (i + 4 <= len) {
        if (data[i] != 0xFF) bad("marker expected");
        const int m = data[i + 1];
        if (m == 0xFF) {
            i++;
            continue;
        }
        if (m == 0xD9) break;
        const size_t L = be16(data + i + 2);
        if (L < 2 || i + 2 + L > len) bad("truncated segment");
        const uint8_t *seg = data + i + 4;
        const size_t n = L - 2;
        if (m == 0xC0 || m == 0xC1) {
            if (n < 6) bad("short SOF");
            if (seg[0] != 8) unsupported("sample precision other than 8 bits");
            h.height = be16(seg + 1), h.width = be16(seg + 3), h.ncomp = seg[5];
            if (h.width <= 0 || h.height <= 0) bad("empty image");
            if (h.ncomp != 1 && h.ncomp != 3) unsupported("with " + std::to_string(h.ncomp) + " components");
            if (n < (size_t)6 + 3 * h.ncomp) bad("short SOF");
            for (int c = 0; c < h.ncomp; c++) {
                h.comp_id[c] = seg[6 + 3 * c];
                h.hs[c] = seg[7 + 3 * c] >> 4, h.vs[c] = seg[7 + 3 * c] & 15;
                h.tq[c] = seg[8 + 3 * c];
                if (h.tq[c] > 3) bad("quantisation table index");
            }
            h.have_sof = true;
        } else if (m == 0xC2) {
            unsupported("progressive (SOF2): MJPG streams are baseline");
        } else if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            unsupported("coding process SOF" + std::to_string(m - 0xC0));
        } else if (m == 0xDB) {
            size_t o = 0;
            while (o < n) {
                const int pq = seg[o] >> 4, tq = seg[o] & 15;
                if (tq > 3) bad("quantisation table index");
                if (pq != 0) unsupported("16-bit quantisation tables");
                if (o + 65 > n) bad("short DQT");
                for (int k = 0; k < 64; k++) h.qt[tq][ZIGZAG[k]] = seg[o + 1 + k];   // natural (row-major) order
                h.have_qt[tq] = true;
                o += 65;
            }
        } else if (m == 0xC4) {
            size_t o = 0;
            while (o < n) {
                if (o + 17 > n) bad("short DHT");
                const int tc = seg[o] >> 4, th = seg[o] & 15;
                if (tc > 1 || th > 3) bad("Huffman table index");
                int total = 0;
                for (int k = 0; k < 16; k++) total += seg[o + 1 + k];
                if (total > 256 || o + 17 + total > n) bad("short DHT");
                JpegHeader::RawHuff &r = h.huff[tc][th];
                memcpy(r.counts, seg + o + 1, 16);
                memcpy(r.symbols, seg + o + 17, total);
                r.nsym = total;
                r.present = true;
                o += 17 + total;
            }
        } else if (m == 0xDD) {
            if (n < 2) bad("short DRI");
            h.restart_interval = be16(seg);
        } else if (m == 0xDA) {
            if (!h.have_sof) bad("SOS before SOF");
            if (n < 1 || seg[0] != h.ncomp) unsupported("non-interleaved scans");
            if (n < (size_t)1 + 2 * h.ncomp + 3) bad("short SOS");
            for (int c = 0; c < h.ncomp; c++) {
                if (seg[1 + 2 * c] != h.comp_id[c]) bad("scan component order");
                h.td[c] = seg[2 + 2 * c] >> 4, h.ta[c] = seg[2 + 2 * c] & 15;
            }
            h.scan_offset = i + 2 + L;
            break;
        }
        i += 2 + L;
    }
    if (!h.have_sof || h.scan_offset == 0) bad("no baseline frame / scan found");
    h.hmax = h.vmax = 1;
    for (int c = 0; c < h.ncomp; c++) h.hmax = std::max(h.hmax, h.hs[c]), h.vmax = std::max(h.vmax, h.vs[c]);
    if (h.ncomp == 1) h.hs[0] = h.vs[0] = h.hmax = h.vmax = 1;     // a single-component scan is never interleaved
    if (h.ncomp == 3) {
        const bool ok = h.hs[1] == 1 && h.vs[1] == 1 && h.hs[2] == 1 && h.vs[2] == 1 &&
                        ((h.hs[0] == 1 && h.vs[0] == 1) || (h.hs[0] == 2 && h.vs[0] == 1) || (h.hs[0] == 2 && h.vs[0] == 2));
        if (!ok) unsupported("sampling factors other than 4:4:4 / 4:2:2 / 4:2:0");
    }
    h.mcus_x = (h.width + 8 * h.hmax - 1) / (8 * h.hmax);
    h.mcus_y = (h.height + 8 * h.vmax - 1) / (8 * h.vmax);
    for (int c = 0; c < h.ncomp; c++) {
        if (!h.have_qt[h.tq[c]]) bad("missing quantisation table");
        if (!h.huff[0][h.td[c]].present || !h.huff[1][h.ta[c]].present) bad("missing Huffman table");
        h.blocks_w[c] = h.mcus_x * h.hs[c];
        h.blocks_h[c] = h.mcus_y * h.vs[c];
    }
    return h;
}

// coeffs[c]: blocks_h[c] * blocks_w[c] blocks of 64 int16 in natural order (caller-allocated).
void jpeg_decode_coefficients(const uint8_t *data, size_t len, const JpegHeader &h, int16_t *const coeffs[3]) {
    HuffTable dc[4], ac[4];
    for (int t = 0; t < 4; t++) {
        if (h.huff[0][t].present) dc[t].build(h.huff[0][t].counts, h.huff[0][t].symbols, h.huff[0][t].nsym);
        if (h.huff[1][t].present) ac[t].build(h.huff[1][t].counts, h.huff[1][t].symbols, h.huff[1][t].nsym);
    }
    BitReader br(data + h.scan_offset, data + len);
    int pred[3] = {0, 0, 0};
    int restart_left = h.restart_interval;
    int next_rst = 0;
    for (int my = 0; my < h.mcus_y; my++) {
        for (int mx = 0; mx < h.mcus_x; mx++) {
            if (h.restart_interval && restart_left == 0) {
                // byte-align, expect RSTn
                const uint8_t *q = br.p;
                while (q + 1 < data + len && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) q++;
                if (q + 1 >= data + len) bad("missing restart marker");
                if (q[1] != 0xD0 + next_rst) bad("restart marker out of sequence");
                next_rst = (next_rst + 1) & 7;
                br.reset_at(q + 2);
                pred[0] = pred[1] = pred[2] = 0;
                restart_left = h.restart_interval;
            }
            for (int c = 0; c < h.ncomp; c++) {
                const HuffTable &tdc = dc[h.td[c]], &tac = ac[h.ta[c]];
                for (int by = 0; by < h.vs[c]; by++)
                    for (int bx = 0; bx < h.hs[c]; bx++) {
                        int16_t *blk = coeffs[c] + ((size_t)(my * h.vs[c] + by) * h.blocks_w[c] + (mx * h.hs[c] + bx)) * 64;
                        memset(blk, 0, 64 * sizeof(int16_t));
                        const int s = decode_symbol(br, tdc);
                        if (s > 15) bad("DC category");
                        int diff = 0;
                        if (s) diff = extend(br.get(s), s);
                        pred[c] += diff;
                        blk[0] = (int16_t)pred[c];
                        for (int k = 1; k < 64;) {
                            const int rs = decode_symbol(br, tac);
                            const int r = rs >> 4, sz = rs & 15;
                            if (sz == 0) {
                                if (r != 15) break;        // EOB
                                k += 16;                   // ZRL
                                continue;
                            }
                            k += r;
                            if (k > 63) bad("AC coefficient index");
                            blk[ZIGZAG[k]] = (int16_t)extend(br.get(sz), sz);
                            k++;
                        }
                    }
            }
            if (h.restart_interval) restart_left--;
        }
    }
}

}  // namespace zb
