// jb_bits.h -- bit-level helpers of the run-length / byte-stream stages.
//
// Everything here is plain C++ marked JB_HD so that the same code runs inside the
// CUDA kernels and in the host-side unit harness (csrc/host_check.cpp).
//
// Stream format (reference pipeline/rle_byte_stream.py:48-88, util.py:115-131,203-221):
//   per non-zero coefficient at zigzag position i with previous non-zero p:
//     run = i - p - 1; (run / 15) times the byte 0xF0 ("15 zeros", util.py:149-154);
//     then run % 15 in 4 bits, size in 4 bits, and `size` amplitude bits: a sign bit
//     (1 = positive, util.py:120-123) followed by |amp| in size-1 bits, where
//     size = bit_length(|amp|) + 1 (util.py:157);
//   a block ends with eight zero bits (EOB) and is zero-padded to a byte boundary
//     (rle_byte_stream.py:55-56,70-72); bits are MSB first.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define JB_HD __host__ __device__ __forceinline__
#else
#define JB_HD inline
#endif

#define JB_MAX_RUN 15
#define JB_MAX_AMP 16383          // size = bit_length + 1 must stay <= 15 (util.py:170-171)

JB_HD uint32_t jb_bswap32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0, 0x0123);
#else
    return __builtin_bswap32(x);
#endif
}

JB_HD int jb_bitlen(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return 32 - __clz((int)x);
#else
    return x ? 32 - __builtin_clz(x) : 0;
#endif
}

// worst-case bytes of one block: every coefficient (run 0, size 15) + EOB
JB_HD int jb_max_block_bytes(int n) { return (23 * n + 8 + 7) / 8; }

// ---------------------------------------------------------------------------------
// MSB-first bit writer into big-endian 32-bit words (so that the byte order in
// memory is the stream order).  `out` must be 4-byte aligned and hold
// ceil(max_block_bytes / 4) words.
// ---------------------------------------------------------------------------------
struct JbBitWriter {
    uint32_t* out;
    uint64_t acc;     // low `nacc` bits are pending
    int nacc;         // < 32 between calls
    int nwords;       // words produced so far (stored only while < cap)
    int cap;          // capacity of `out` in words; bits beyond it are counted, not stored

    JB_HD void init(uint32_t* o, int cap_words = 0x7fffffff) { out = o; acc = 0; nacc = 0; nwords = 0; cap = cap_words; }

    JB_HD void put(uint32_t v, int k) {          // k <= 23
        acc = (acc << k) | (uint64_t)v;
        nacc += k;
        if (nacc >= 32) {
            nacc -= 32;
            if (nwords < cap) out[nwords] = jb_bswap32((uint32_t)(acc >> nacc));
            ++nwords;
        }
    }

    // returns the block length in bytes (after zero padding to a byte boundary)
    JB_HD uint32_t finish() {
        uint32_t bits = (uint32_t)nwords * 32u + (uint32_t)nacc;
        if (nacc && nwords < cap) out[nwords] = jb_bswap32((uint32_t)(acc << (32 - nacc)));
        return (bits + 7u) >> 3;
    }
};

// Same interface, writing bytes to an arbitrary (unaligned) address: the slow path for
// blocks that do not fit the fixed-size staging row of the specialised kernel.
struct JbByteWriter {
    uint8_t* out;
    uint64_t acc;
    int nacc;         // < 8 between calls
    uint32_t nbytes;

    JB_HD void init(uint8_t* o) { out = o; acc = 0; nacc = 0; nbytes = 0; }
    JB_HD void put(uint32_t v, int k) {
        acc = (acc << k) | (uint64_t)v;
        nacc += k;
        while (nacc >= 8) { nacc -= 8; out[nbytes++] = (uint8_t)(acc >> nacc); }
    }
    JB_HD uint32_t finish() {
        if (nacc) { out[nbytes++] = (uint8_t)(acc << (8 - nacc)); nacc = 0; }
        return nbytes;
    }
};

// Emit the code(s) for one non-zero coefficient.  Returns false if the amplitude
// does not fit (BadRleCodeError in the reference); nothing is emitted then.
template <typename Writer>
JB_HD bool jb_put_coefficient(Writer& w, int run, int amp) {
    uint32_t mag = (uint32_t)(amp < 0 ? -amp : amp);
    if (mag > JB_MAX_AMP) return false;
    while (run >= JB_MAX_RUN) { w.put(0xF0u, 8); run -= JB_MAX_RUN; }
    int size = jb_bitlen(mag) + 1;
    uint32_t code = ((((uint32_t)run << 4) | (uint32_t)size) << size)
                  | ((amp > 0 ? 1u : 0u) << (size - 1)) | mag;
    w.put(code, 8 + size);
    return true;
}

// Run-length encode + pack one block of n zigzag-ordered coefficients (generic n).
// Returns the byte length; *bad_pos receives the zigzag position of the first
// coefficient whose amplitude does not fit (or -1), *bad_run its run length.
template <typename CoefT>
JB_HD uint32_t jb_pack_block(const CoefT* c, int n, uint32_t* out, int* bad_pos, int* bad_run) {
    JbBitWriter w;
    w.init(out);
    int prev = -1;
    *bad_pos = -1;
    *bad_run = 0;
    for (int p = 0; p < n; ++p) {
        int a = (int)c[p];
        if (a == 0) continue;
        int run = p - prev - 1;
        prev = p;
        if (!jb_put_coefficient(w, run, a) && *bad_pos < 0) { *bad_pos = p; *bad_run = run % JB_MAX_RUN; }
    }
    w.put(0u, 8);                                   // EOB
    return w.finish();
}

// ---------------------------------------------------------------------------------
// MSB-first bit reader over a byte array.  Bytes at and after `limit` read as zero;
// consuming bits beyond `limit` marks the reader as overrun.
// ---------------------------------------------------------------------------------
struct JbBitReader {
    const uint8_t* base;
    uint32_t pos;      // bit position
    uint32_t limit;    // bytes readable
    bool overrun;

    JB_HD void init(const uint8_t* b, uint32_t byte_start, uint32_t byte_limit) {
        base = b; pos = byte_start * 8u; limit = byte_limit; overrun = false;
    }
    JB_HD uint32_t byte_at(uint32_t i) const { return i < limit ? (uint32_t)base[i] : 0u; }
    // k <= 16 bits; reading past `limit` bytes yields zeros and sets `overrun`
    JB_HD uint32_t get(int k) {
        uint32_t byte = pos >> 3;
        uint32_t window = (byte_at(byte) << 16) | (byte_at(byte + 1) << 8) | byte_at(byte + 2);
        uint32_t v = (window >> (24 - (pos & 7u) - k)) & ((1u << k) - 1u);
        pos += (uint32_t)k;
        if (pos > limit * 8u) overrun = true;
        return v;
    }
};

#define JB_PARSE_OK 0
#define JB_PARSE_BAD 1      // invalid code, too many coefficients, or ran past the data

// Walk one block starting at byte `start` (rle_byte_stream.py:74-88) without
// decoding amplitudes.  On success *end_byte is the first byte after the block
// (its padding skipped).  n = coefficients per block; blocks longer than
// max_bytes are rejected.
JB_HD int jb_parse_block_extent(const uint8_t* data, uint32_t start, uint32_t limit, int n,
                                uint32_t max_bytes, uint32_t* end_byte) {
    JbBitReader r;
    r.init(data, start, limit);
    int count = 0;
    const uint32_t bit_limit = (start + max_bytes) * 8u;
    for (;;) {
        if (r.pos + 8u > bit_limit) return JB_PARSE_BAD;
        uint32_t head = r.get(8);
        if (r.overrun) return JB_PARSE_BAD;
        uint32_t run = head >> 4, size = head & 15u;
        if (size == 0u) {
            if (run == 0u) break;                         // EOB
            if (run != (uint32_t)JB_MAX_RUN) return JB_PARSE_BAD;   // (r,0,0), 0<r<15: util.py:176-177
            count += JB_MAX_RUN;
        } else {
            if (size == 1u) return JB_PARSE_BAD;          // the reference's int('', 2) fails
            count += (int)run + 1;
            r.pos += size;
        }
        if (count > n) return JB_PARSE_BAD;
    }
    uint32_t e = (r.pos + 7u) >> 3;
    if (e > limit || e - start > max_bytes) return JB_PARSE_BAD;
    *end_byte = e;
    return JB_PARSE_OK;
}

// Decode one block into n coefficients at out[perm[i]] (perm maps zigzag position ->
// storage index; pass nullptr for identity).  `out` must be zero-filled by the caller.
template <typename CoefT, typename PermT>
JB_HD int jb_decode_block(const uint8_t* data, uint32_t start, uint32_t limit, int n,
                          CoefT* out, const PermT* perm) {
    JbBitReader r;
    r.init(data, start, limit);
    int count = 0;
    for (;;) {
        uint32_t head = r.get(8);
        if (r.overrun) return JB_PARSE_BAD;
        uint32_t run = head >> 4, size = head & 15u;
        if (size == 0u) {
            if (run == 0u) break;
            if (run != (uint32_t)JB_MAX_RUN) return JB_PARSE_BAD;
            count += JB_MAX_RUN;
            if (count > n) return JB_PARSE_BAD;
        } else {
            if (size == 1u) return JB_PARSE_BAD;
            count += (int)run;
            if (count >= n) return JB_PARSE_BAD;
            uint32_t raw = r.get((int)size);
            if (r.overrun) return JB_PARSE_BAD;
            int mag = (int)(raw & ((1u << (size - 1)) - 1u));
            int amp = (raw >> (size - 1)) ? mag : -mag;    // sign bit 1 = positive
            out[perm ? (int)perm[count] : count] = (CoefT)amp;
            count += 1;
        }
    }
    return JB_PARSE_OK;
}
