// Baseline JPEG front end (host side): see jpeg_host.cpp.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace zb {

struct JpegHeader {
    int width = 0, height = 0, ncomp = 0;
    int comp_id[3] = {0, 0, 0}, hs[3] = {1, 1, 1}, vs[3] = {1, 1, 1}, tq[3] = {0, 0, 0}, td[3] = {0, 0, 0}, ta[3] = {0, 0, 0};
    int hmax = 1, vmax = 1, mcus_x = 0, mcus_y = 0;
    int blocks_w[3] = {0, 0, 0}, blocks_h[3] = {0, 0, 0};    // per component, padded to whole MCUs
    uint16_t qt[4][64] = {};                                  // natural (row-major) order
    bool have_qt[4] = {false, false, false, false};
    struct RawHuff {
        bool present = false;
        uint8_t counts[16] = {};
        uint8_t symbols[256] = {};
        int nsym = 0;
    } huff[2][4];                                             // [dc / ac][table id]
    int restart_interval = 0;
    bool have_sof = false;
    size_t scan_offset = 0;                                   // first byte of the entropy-coded segment
};

// Throws std::runtime_error: "jpeg: ..." for malformed data, "unsupported op: JPEG ..." for progressive / CMYK / 12-bit ...
JpegHeader jpeg_parse_header(const uint8_t *data, size_t len);
// Entropy decoding into a sparse coefficient stream: block b (component-major, row-major inside a component) owns
// stream[start[b] .. start[b] + 3 * count[b]): {natural-order index, value lo, value hi} per non-zero coefficient.
// start / count: total_blocks(h) entries, caller-allocated.
void jpeg_decode_sparse(const uint8_t *data, size_t len, const JpegHeader &h, uint32_t *start, uint8_t *count, std::vector<uint8_t> &stream);
inline int jpeg_total_blocks(const JpegHeader &h) {
    int n = 0;
    for (int c = 0; c < h.ncomp; c++) n += h.blocks_w[c] * h.blocks_h[c];
    return n;
}

}  // namespace zb
