// host_check.cpp -- host build of the bit-level helpers in jb_bits.h, so that the exact
// code the kernels run per thread (RLE packing, block extent parse, block decode) can be
// checked against the oracle on a machine without a GPU (tests/test_host_bits.py).
// Built as libjbhostcheck.so by __graft_entry__.build(); not part of the product path.
#include <stdint.h>
#include <string.h>
#include <vector>

#include "jb_bits.h"

extern "C" {

// Pack nblocks blocks of n int32 zigzag coefficients.  out must hold nblocks * max_block_bytes.
// Returns total bytes, or -(1 + index of first bad block) if an amplitude does not fit.
long long hc_pack_blocks(const int32_t* coefs, int nblocks, int n, uint8_t* out, uint32_t* block_len,
                         int* bad_pos, int* bad_run) {
    std::vector<uint32_t> stage((jb_max_block_bytes(n) + 3) / 4 + 2);
    long long total = 0, bad = 0;
    for (int b = 0; b < nblocks; ++b) {
        int bp, br;
        uint32_t len = jb_pack_block<int32_t>(coefs + (size_t)b * n, n, stage.data(), &bp, &br);
        if (bp >= 0 && bad == 0) { bad = -(1 + b); *bad_pos = bp; *bad_run = br; }
        memcpy(out + total, stage.data(), len);
        block_len[b] = len;
        total += len;
    }
    return bad ? bad : total;
}

// Walk a whole stream block by block with the extent parser; returns the block count, or -1.
long long hc_walk_stream(const uint8_t* data, uint32_t len, int n, uint32_t* starts, long long cap) {
    uint32_t pos = 0;
    long long k = 0;
    const uint32_t maxb = (uint32_t)jb_max_block_bytes(n);
    while (pos < len) {
        uint32_t e;
        if (jb_parse_block_extent(data, pos, len, n, maxb, &e) != JB_PARSE_OK) return -1;
        if (k < cap) starts[k] = pos;
        ++k;
        pos = e;
    }
    return k;
}

// Decode block at `start` into n int32 coefficients (zigzag order).  Returns 0 or 1 (bad).
int hc_decode_block(const uint8_t* data, uint32_t start, uint32_t len, int n, int32_t* out) {
    memset(out, 0, sizeof(int32_t) * n);
    return jb_decode_block<int32_t, uint16_t>(data, start, len, n, out, (const uint16_t*)nullptr);
}

int hc_parse_extent(const uint8_t* data, uint32_t start, uint32_t len, int n, uint32_t* end) {
    return jb_parse_block_extent(data, start, len, n, (uint32_t)jb_max_block_bytes(n), end);
}

}  // extern "C"
