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
#include <vector>

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

// ITU T.81 Annex K.3 - K.6 "typical" Huffman tables: MJPG frames from webcams routinely omit their DHT segments and mean
// these (libjpeg-turbo installs the same defaults, jstdhuff.c).
const uint8_t STD_DC_LUM_BITS[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t STD_DC_CHR_BITS[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t STD_DC_VALS[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t STD_AC_LUM_BITS[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t STD_AC_LUM_VALS[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91,
    0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a,
    0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53,
    0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79,
    0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5,
    0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9,
    0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2,
    0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const uint8_t STD_AC_CHR_BITS[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t STD_AC_CHR_VALS[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14,
    0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17,
    0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a,
    0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78,
    0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7,
    0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2,
    0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

void set_std(JpegHeader::RawHuff &r, const uint8_t *bits, const uint8_t *vals, int n) {
    memcpy(r.counts, bits, 16);
    memcpy(r.symbols, vals, n);
    r.nsym = n;
    r.present = true;
}

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
    // MJPG without DHT: the standard tables (luminance = table 0, chrominance = table 1)
    if (!h.huff[0][0].present) set_std(h.huff[0][0], STD_DC_LUM_BITS, STD_DC_VALS, 12);
    if (!h.huff[0][1].present) set_std(h.huff[0][1], STD_DC_CHR_BITS, STD_DC_VALS, 12);
    if (!h.huff[1][0].present) set_std(h.huff[1][0], STD_AC_LUM_BITS, STD_AC_LUM_VALS, 162);
    if (!h.huff[1][1].present) set_std(h.huff[1][1], STD_AC_CHR_BITS, STD_AC_CHR_VALS, 162);
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

// Entropy decoding into a SPARSE coefficient stream (most quantised coefficients are zero, and the stream is what crosses
// PCIe): for flat block index b = block_base[component] + by * blocks_w + bx, `start[b]` / `count[b]` delimit its non-zero
// coefficients in `stream`, 3 bytes each: natural-order index, value low byte, value high byte.
void jpeg_decode_sparse(const uint8_t *data, size_t len, const JpegHeader &h, uint32_t *start, uint8_t *count, std::vector<uint8_t> &stream) {
    HuffTable dc[4], ac[4];
    for (int t = 0; t < 4; t++) {
        if (h.huff[0][t].present) dc[t].build(h.huff[0][t].counts, h.huff[0][t].symbols, h.huff[0][t].nsym);
        if (h.huff[1][t].present) ac[t].build(h.huff[1][t].counts, h.huff[1][t].symbols, h.huff[1][t].nsym);
    }
    int block_base[3] = {0, 0, 0};
    for (int c = 1; c < h.ncomp; c++) block_base[c] = block_base[c - 1] + h.blocks_w[c - 1] * h.blocks_h[c - 1];
    // every non-zero coefficient costs at least 3 bits of entropy-coded data (a >= 2-bit code and >= 1 value bit), so
    // 3 bytes per coefficient never exceed 8 x the compressed size (+ one triple per block for the DC terms)
    stream.resize((len - h.scan_offset) * 8 + (size_t)3 * (h.mcus_x * h.mcus_y * 6) + 64);
    uint8_t *wp = stream.data();
    BitReader br(data + h.scan_offset, data + len);
    int pred[3] = {0, 0, 0};
    int restart_left = h.restart_interval;
    int next_rst = 0;
    auto put = [&](int idx, int v) {
        wp[0] = (uint8_t)idx, wp[1] = (uint8_t)(v & 255), wp[2] = (uint8_t)((v >> 8) & 255);
        wp += 3;
    };
    for (int my = 0; my < h.mcus_y; my++) {
        for (int mx = 0; mx < h.mcus_x; mx++) {
            if (h.restart_interval && restart_left == 0) {
                const uint8_t *q = br.p;      // the reader never consumes a marker: the RSTn is at or after its position
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
                        const int b = block_base[c] + (my * h.vs[c] + by) * h.blocks_w[c] + (mx * h.hs[c] + bx);
                        start[b] = (uint32_t)(wp - stream.data());
                        if ((size_t)(wp - stream.data()) + 3 * 64 > stream.size()) bad("coefficient stream overflow");
                        int n = 0;
                        const int sdc = decode_symbol(br, tdc);
                        if (sdc > 15) bad("DC category");
                        if (sdc) pred[c] += extend(br.get(sdc), sdc);
                        if (pred[c]) put(0, pred[c]), n++;
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
                            put(ZIGZAG[k], extend(br.get(sz), sz));
                            n++;
                            k++;
                        }
                        count[b] = (uint8_t)n;
                    }
            }
            if (h.restart_interval) restart_left--;
        }
    }
    stream.resize((size_t)(wp - stream.data()));
}

}  // namespace zb
