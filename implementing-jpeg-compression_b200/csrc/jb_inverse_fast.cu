// jb_inverse_fast.cu -- decompress direction, specialised for dct_size 8 / block_size 4.
//
// One warp owns a chunk of 32 consecutive blocks of one plane (block offsets come from
// jb_framing.cu):
//   * the chunk's bytes (contiguous in the stream) are copied to shared memory, then lane t
//     decodes block t (rle_byte_stream.py:74-88, run_length_encoding.py:31-41) straight into
//     natural (un-zigzagged, zigzag_order.py:101-119) order as int16;
//   * 8 iterations over 4 blocks: lane (u, b) dequantises row u (quantizers.py restore, with
//     the 1/|c_k|^2 scale of transforms.py:14-26 folded in), 8-point inverse transform along
//     the row, transpose through shared memory, 8-point inverse transform down the column,
//     np.round + clamp (basis_change.py:43, normalization.py:10-14);
//   * default (ROWS): the 8-bit samples of the chunk stay in shared memory and the chunk goes out
//     pixel row by pixel row -- a lane pair owns a block, one 128-bit store instruction covers 512
//     contiguous bytes, the 4x4 replication (util.inflate) happens in registers, the crops of
//     dct_padding.py:11-21 / padding.py:14-16 by predication.  HBM takes the chunk's 1 KB row runs at
//     6.1 TB/s, 32-row x 128-byte tiles only at 5.2 (tools/dram_probe_rows.cu);
//   * JB_FLAG_TILE_DECODER (the former default, kept for comparison): 4x4 replication into a 32-row x
//     128-byte tile in shared memory, which goes out by TMA store (cp.async.bulk.tensor shared ->
//     global, clipping at the tensor bounds performs the crops); tiles that wrap a block row or belong
//     to a partial group use bounded stores;
//   * chunks are claimed from a device-wide counter (see the kernel).
#include <cuda.h>
#include <string.h>

#include "jb_common.cuh"
#include "jb_fast_common.cuh"
#include "jb_inverse.cuh"

#define FI_WARPS 16
#define FI_RING 2
#define FI_SAMPLE_PITCH 68              // bytes per block in the sample buffer of the row-store variant (64 + 4:
                                        // the word reads of a store row fall into distinct banks)
#define FI_STREAM_WORDS 400             // staged chunk bytes (1.6 KB; the average chunk is ~0.6 KB); denser
                                        // chunks are decoded straight from global memory

// Row-store variant: 72 registers and 7.8 KB of shared memory per warp would allow 28 warps per SM, but the kernel is
// bound by the HBM write path, not by latency: measured 1.12 / 1.10 / 1.12 / 1.14 / 1.06 / 1.11 / 1.13 / 1.15 ms with
// 8 / 10 / 12 / 14 / 16 / 20 / 24 / 28 warps (1024 x 1080p, profiles/r1_sweep_inv_warps.txt).
#define FI_ROWS_MAX_WARPS 16
#define FI_ROWS_WARPS 16
#define FI_ROWS_STAGE_WORDS 512

// per-warp shared memory of the row-store variant: the chunk's bytes, then (once decoded) its 8-bit samples
struct __align__(128) FiRowsSmem {
    uint8_t buf[2304];                  // max(4 FI_ROWS_STAGE_WORDS, 32 FI_SAMPLE_PITCH)
    float scr[4 * FF_BLK_W];
    uint32_t coef[JB_CHUNK * FF_COEF_W];
};

struct __align__(128) FiWarpSmem {
    uint8_t tile[FI_RING][FF_TILE_BYTES];
    float scr[4 * FF_BLK_W];
    uint32_t coef[JB_CHUNK * FF_COEF_W];
};

struct FiKernelArgs {
    JbInvArgs a;
    int use_tma;
    int aligned;
};

// Decode one block from 32-bit words (staged in shared memory, or in global memory for chunks too
// dense to stage) into natural order.  `bitpos` / `bitlimit` count from words[0].  Returns 0 or 1
// (malformed).
template <bool GLOBAL>
__device__ __forceinline__ int fi_decode_block(const uint32_t* words, uint32_t nwords, uint32_t bitpos,
                                               uint32_t bitlimit, int16_t* row, const uint8_t* izz) {
    uint32_t widx = bitpos >> 5;
    auto ld = [&](uint32_t i) -> uint32_t {
        if (i >= nwords) return 0u;
        return jb_bswap32(GLOBAL ? __ldg(words + i) : words[i]);
    };
    int nb = 32 - (int)(bitpos & 31u);
    uint64_t buf = ld(widx++);
    uint32_t used = bitpos;
    int count = 0;
    for (;;) {
        if (nb < 23) { buf = (buf << 32) | ld(widx++); nb += 32; }
        const uint32_t head = (uint32_t)(buf >> (nb - 8)) & 0xFFu;
        const uint32_t run = head >> 4, size = head & 15u;
        if (size == 0u) {
            nb -= 8; used += 8;
            if (run == 0u) return used > bitlimit ? 1 : 0;            // EOB
            if (run != (uint32_t)JB_MAX_RUN) return 1;
            count += JB_MAX_RUN;
            if (count > 64 || used > bitlimit) return 1;
            continue;
        }
        if (size == 1u) return 1;
        count += (int)run;
        if (count >= 64) return 1;
        const uint32_t raw = (uint32_t)(buf >> (nb - 8 - (int)size)) & ((1u << size) - 1u);
        nb -= 8 + (int)size; used += 8u + size;
        if (used > bitlimit) return 1;
        const int mag = (int)(raw & ((1u << (size - 1)) - 1u));
        row[izz[count]] = (int16_t)((raw >> (size - 1)) ? mag : -mag);
        ++count;
    }
}

struct FiCursor { int it, by, bx; };

// kind 0: 4 blocks of one block row, all inside the image -> TMA store or STG.128
// kind 1: same, but crossing the right / bottom edge     -> TMA store (clips) or bounded stores
// kind 2: block-row wrap, partial group, unaligned plane -> bounded stores
__device__ __forceinline__ int fi_tile_kind(const JbGeom& g, int nvalid, const FiCursor& c, bool aligned) {
    if (!aligned || 4 * c.it + 4 > nvalid || c.bx + 4 > g.hb) return 2;
    return ((c.bx + 4) * 32 <= g.W && (c.by + 1) * 32 <= g.H) ? 0 : 1;
}

template <bool DFT, int MODE, bool ROWS>
__global__ void __launch_bounds__((ROWS ? FI_ROWS_MAX_WARPS : FI_WARPS) * 32, 1)
jb_inv_fast_kernel(const __grid_constant__ CUtensorMap tmap, const FiKernelArgs ka) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const JbInvArgs& a = ka.a;
    const JbGeom& g = a.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* s_izz = (uint8_t*)smem_raw;                               // zigzag position -> natural index
    const int NWARPS = blockDim.x >> 5;
    constexpr unsigned STAGE_WORDS = ROWS ? FI_ROWS_STAGE_WORDS : FI_STREAM_WORDS;
    FiWarpSmem& ws = *(FiWarpSmem*)(smem_raw + 128 + (size_t)warp * sizeof(FiWarpSmem));           // tile variant
    FiRowsSmem& wr = *(FiRowsSmem*)(smem_raw + 128 + (size_t)warp * sizeof(FiRowsSmem));           // row variant
    uint32_t* const w_coef = ROWS ? wr.coef : ws.coef;
    float* const w_scr = ROWS ? wr.scr : ws.scr;
    jb_pdl_trigger();
    // the tables are constant once built (JB_FLAG_REUSE_TABLES): their loads may overlap the kernel before this one
    const bool tables_const = (g.flags & JB_FLAG_REUSE_TABLES) != 0;
    if (!tables_const) jb_pdl_wait();
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_izz[i] = (uint8_t)a.t.izz[i];
    __syncthreads();

    const int li = lane >> 2, lb = lane & 3;
    const int ncol = DFT ? (li < 5 ? li : 12 - li) : li;               // sample column after the column stage
    // dequantiser of row u = li with the inverse-transform scale folded in (powers of two: exact)
    float dq[8];
    #pragma unroll
    for (int v = 0; v < 8; ++v) {
        const float sc = DFT ? (1.0f / 64.0f) : ((li == 0 ? 0.125f : 0.25f) * (v == 0 ? 0.125f : 0.25f));
        dq[v] = a.t.dqmult[li * 8 + v] * sc;
    }
    if (tables_const) jb_pdl_wait();                                   // from here on: data the kernels before this one wrote

    // Chunks are claimed from a device-wide counter: with a fixed deal the slowest SM finished 36 % later than the
    // fastest (ncu, sm__cycles_active min / max), i.e. the kernel ran 13 % longer than its average SM.  Without a counter
    // (stage entry points, whose workspace holds only the tables) chunks are dealt round-robin.
    // A chunk begins with three dependent round trips -- its ticket, the offsets of its blocks, its bytes -- which a
    // warp on its own would sit out (a quarter of the kernel's stall samples).  They are spread over the chunk before:
    // the ticket of chunk n + 1 is drawn at the top of chunk n and picked up behind its entropy decode, where the loads
    // of the new chunk's block offsets are issued; a few transform iterations later those have landed and the first
    // FI_PRE_WORDS words per lane of its bytes are fetched into registers (the shared-memory stage still holds chunk n).
    // No deeper: a warp that holds tickets further ahead widens the set of chunks in flight across the GPU, and the
    // pixel rows of neighbouring chunks stop meeting in the same DRAM pages (measured with a pipeline two chunks deep:
    // 1.066 -> 1.123 ms at 1024 frames).  The first chunk of every warp is dealt statically, CTA-major (a single frame
    // spreads over all SMs).
    const unsigned total_warps = gridDim.x * NWARPS;
    unsigned store_seq = 0;
    constexpr int FI_PRE_WORDS = 8;                                     // 1 KB per chunk: an average chunk is ~0.6 KB
    struct FiChunk {
        int plane, blk0, nvalid;
        unsigned my_start, end_raw;                                     // block offsets as loaded (lane < nvalid; lane 0)
        unsigned long long len, off;
    };
    auto describe = [&](unsigned c, FiChunk& k) {                       // issues the loads of a chunk's offsets
        k.plane = (int)(c / (unsigned)g.cpp);
        k.blk0 = (int)(c % (unsigned)g.cpp) * JB_CHUNK;
        k.nvalid = jb_min(JB_CHUNK, g.nblocks - k.blk0);
        k.my_start = 0u; k.end_raw = 0u; k.len = 0ull; k.off = 0ull;
        if (MODE != 2) {
            k.len = a.plane_len[k.plane];
            k.off = a.plane_off[k.plane];
            const unsigned* bs = a.block_start + (size_t)k.plane * g.nblocks + k.blk0;
            if (lane < k.nvalid) k.my_start = bs[lane];
            if (lane == 0 && k.blk0 + k.nvalid < g.nblocks) k.end_raw = bs[k.nvalid];
        }
    };
    struct FiBytes { bool ok; unsigned c_start, c_end, mis, nwords; const uint32_t* wsrc; };
    auto locate = [&](const FiChunk& k, FiBytes& y) {                   // (consumes the loads of describe)
        y.c_start = __shfl_sync(0xffffffffu, k.my_start, 0);
        const unsigned e0 = __shfl_sync(0xffffffffu, k.end_raw, 0);
        y.c_end = (k.blk0 + k.nvalid < g.nblocks) ? e0 : (unsigned)k.len;
        // (a stream must lie inside the caller's buffer: device-supplied offsets are not trusted)
        y.ok = y.c_start <= y.c_end && y.c_end <= k.len && k.off <= a.in_bytes && k.len <= a.in_bytes - k.off;
        const unsigned long long addr0 = (unsigned long long)(uintptr_t)(a.in + k.off + y.c_start);
        y.mis = (unsigned)(addr0 & 3ull);
        y.wsrc = (const uint32_t*)(uintptr_t)(addr0 - y.mis);
        y.nwords = y.ok ? (((y.c_end - y.c_start) + y.mis + 3u) >> 2) : 0u;
    };
    unsigned chunk = blockIdx.x + gridDim.x * (unsigned)warp;
    FiChunk cur_k, next_k;
    uint32_t pre[FI_PRE_WORDS];
    bool pre_valid = false;                                             // pre[] holds the first words of `chunk`
    if (chunk < a.n_chunks) describe(chunk, cur_k);
    while (chunk < a.n_chunks) {
        unsigned pend = 0;                                              // lane 0: the next chunk
        if (a.ticket != nullptr && lane == 0) pend = total_warps + atomicAdd(a.ticket, 1u);
        unsigned next_chunk = 0xFFFFFFFFu;
        bool have_next = false, next_pre = false, next_located = false;
        const int plane = cur_k.plane, blk0 = cur_k.blk0, nvalid = cur_k.nvalid;
        const int nit = (nvalid + 3) >> 2;

        // ---- coefficients of the chunk in natural order ----
        {
            uint4* z = (uint4*)w_coef;
            for (int i = lane; i < JB_CHUNK * FF_COEF_W / 4; i += 32) z[i] = make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
        if (MODE == 2) {
            const int16_t* src = a.coeffs_in + ((size_t)plane * g.nblocks + blk0) * 64;
            for (int idx = lane; idx < nvalid * 64; idx += 32)
                ((int16_t*)(w_coef + (idx >> 6) * FF_COEF_W))[s_izz[idx & 63]] = src[idx];
        } else {
            FiBytes y;
            locate(cur_k, y);
            const unsigned my_start = cur_k.my_start, c_start = y.c_start, c_end = y.c_end, mis = y.mis, nwords = y.nwords;
            const bool ok = y.ok;
            const uint32_t* wsrc = y.wsrc;
            const bool staged = nwords <= STAGE_WORDS;
            // the chunk's bytes are staged in the ring slot that the next tile will use: the store issued
            // from it two tiles ago has to be done reading it
            uint32_t* sbytes = ROWS ? (uint32_t*)wr.buf : (uint32_t*)ws.tile[store_seq % FI_RING];
            if (!ROWS) {
                if (lane == 0) ff_bulk_wait_read<FI_RING - 1>();
                __syncwarp();
            }
            if (ok && staged) {
                if (pre_valid) {
                    #pragma unroll
                    for (int q = 0; q < FI_PRE_WORDS; ++q)
                        if ((unsigned)(lane + 32 * q) < nwords) sbytes[lane + 32 * q] = pre[q];
                    for (unsigned i = lane + 32 * FI_PRE_WORDS; i < nwords; i += 32) sbytes[i] = __ldg(wsrc + i);
                } else {
                    for (unsigned i = lane; i < nwords; i += 32) sbytes[i] = __ldg(wsrc + i);
                }
            }
            __syncwarp();
            int bad = ok ? 0 : 1;
            if (ok && lane < nvalid) {
                int16_t* row = (int16_t*)(w_coef + lane * FF_COEF_W);
                const uint32_t bit0 = (my_start - c_start + mis) * 8u, bitl = (c_end - c_start + mis) * 8u;
                if (my_start < c_start || my_start >= c_end) bad = 1;
                else if (staged) bad = fi_decode_block<false>(sbytes, nwords, bit0, bitl, row, s_izz);
                else bad = fi_decode_block<true>(wsrc, nwords, bit0, bitl, row, s_izz);
            }
            if (__any_sync(0xffffffffu, bad) && lane == 0) jb_set_error(a.status, JB_ERR_BAD_STREAM);
        }
        // the ticket has landed by now: the loads of the next chunk's block offsets go out
        next_chunk = a.ticket != nullptr ? __shfl_sync(0xffffffffu, pend, 0) : chunk + total_warps;
        have_next = next_chunk < a.n_chunks;
        if (have_next) describe(next_chunk, next_k);
        auto prefetch_next = [&]() {                                    // (once per chunk, when the offsets are there)
            next_located = true;
            if (MODE == 2 || !have_next) return;
            FiBytes ny;
            locate(next_k, ny);
            if (ny.ok && ny.nwords <= STAGE_WORDS) {
                #pragma unroll
                for (int q = 0; q < FI_PRE_WORDS; ++q)
                    pre[q] = (unsigned)(lane + 32 * q) < ny.nwords ? __ldg(ny.wsrc + lane + 32 * q) : 0u;
                next_pre = true;
            }
        };
        __syncwarp();

        // ---- 4 blocks per iteration ----
        FiCursor cur;
        cur.it = 0; cur.by = blk0 / g.hb; cur.bx = blk0 - cur.by * g.hb;
        for (int it = 0; it < nit; ++it) {
            if (it == 3) prefetch_next();
            const int slot = (int)(store_seq % FI_RING);
            uint8_t* tile = ws.tile[slot];
            const int kind = ROWS ? 0 : fi_tile_kind(g, nvalid, cur, ka.aligned != 0);
            // the TMA store issued FI_RING tiles ago must have finished reading this slot
            if (!ROWS) {
                if (lane == 0) ff_bulk_wait_read<FI_RING - 1>();
                __syncwarp();
            }

            float z[8], r[8];
            {
                const uint4 w = *(const uint4*)(w_coef + (4 * it + lb) * FF_COEF_W + li * 4);
                z[0] = (float)(short)(w.x & 0xFFFFu) * dq[0]; z[1] = (float)(short)(w.x >> 16) * dq[1];
                z[2] = (float)(short)(w.y & 0xFFFFu) * dq[2]; z[3] = (float)(short)(w.y >> 16) * dq[3];
                z[4] = (float)(short)(w.z & 0xFFFFu) * dq[4]; z[5] = (float)(short)(w.z >> 16) * dq[5];
                z[6] = (float)(short)(w.w & 0xFFFFu) * dq[6]; z[7] = (float)(short)(w.w >> 16) * dq[7];
            }
            if (DFT) ff_rdft8(z, r); else ff_idct8(z, r);
            {
                float4* sc = (float4*)(w_scr + lb * FF_BLK_W + li * 8);
                const int h0 = li & 1;
                const float4 r0 = make_float4(r[0], r[1], r[2], r[3]), r1 = make_float4(r[4], r[5], r[6], r[7]);
                sc[h0] = h0 ? r1 : r0; sc[h0 ^ 1] = h0 ? r0 : r1;
            }
            __syncwarp();
            float col[8], x[8];
            #pragma unroll
            for (int k = 0; k < 8; ++k) col[k] = w_scr[lb * FF_BLK_W + k * 8 + li];
            if (DFT) ff_dft_column_stage(col, li, x); else ff_idct8(col, x);

            // round half-even, clamp to 0..255
            uint32_t* t32 = (uint32_t*)tile;
            if (ROWS) {
                // sample (m, ncol) of block 4 it + lb, one byte each; the chunk goes out row by row below
                uint8_t* sp = wr.buf + (4 * it + lb) * FI_SAMPLE_PITCH + ncol;
                #pragma unroll
                for (int m = 0; m < 8; ++m) {
                    int p = __float_as_int(x[m] + 12582912.0f) - 0x4B400000;
                    sp[m * 8] = (uint8_t)max(0, min(255, p));
                }
                ++cur.it;
                __syncwarp();
                continue;
            }
            // replicate 4x4: word column 8b + ncol, rows 4m..4m+3
            #pragma unroll
            for (int m = 0; m < 8; ++m) {
                int p = __float_as_int(x[m] + 12582912.0f) - 0x4B400000;
                p = max(0, min(255, p));
                const uint32_t wv = (uint32_t)p * 0x01010101u;
                #pragma unroll
                for (int k = 0; k < 4; ++k) t32[(4 * m + k) * 32 + 8 * lb + ncol] = wv;
            }
            __syncwarp();

            const int n0 = blk0 + 4 * it;
            uint8_t* plane_ptr = a.planes_out + (size_t)plane * a.plane_stride;
            if (kind <= 1 && ka.use_tma) {
                if (lane == 0) {
                    ff_fence_proxy_async();
                    ff_tma_store_3d(&tmap, cur.bx * 32, cur.by * 32, plane, tile);
                    ff_bulk_commit();
                }
            } else if (kind == 0) {
                uint8_t* dst = plane_ptr + (size_t)cur.by * 32 * a.row_pitch + (size_t)cur.bx * 32;
                const uint4* t16 = (const uint4*)tile;
                #pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int idx = k * 32 + lane;
                    *((uint4*)(dst + (size_t)(idx >> 3) * a.row_pitch) + (idx & 7)) = t16[idx];
                }
            } else {
                const int nlive = jb_min(4, nvalid - 4 * it);
                for (int idx = lane; idx < 1024; idx += 32) {
                    const int row = idx >> 5, wcol = idx & 31, b = wcol >> 3;
                    if (b >= nlive) continue;
                    const int blk = n0 + b;
                    const int by = blk / g.hb, bx = blk - by * g.hb;
                    const int yy = by * 32 + row, xx = bx * 32 + (wcol & 7) * 4;
                    if (yy >= g.H || xx >= g.W) continue;
                    const uint32_t wv = t32[idx];
                    uint8_t* d = plane_ptr + (size_t)yy * a.row_pitch + xx;
                    const int nbytes = jb_min(4, g.W - xx);
                    for (int q = 0; q < nbytes; ++q) d[q] = (uint8_t)(wv >> (8 * q));
                }
            }
            ++store_seq;
            ++cur.it;
            cur.bx += 4;
            while (cur.bx >= g.hb) { cur.bx -= g.hb; ++cur.by; }
            __syncwarp();
        }
        if (ROWS) {
            // ---- the chunk, pixel row by pixel row: a lane pair owns a block (lane & 1 = its left / right half),
            // blocks blk0 .. blk0+15 in the first store of a row and blk0+16 .. blk0+31 in the second, so one store
            // instruction covers up to 512 contiguous bytes and the chunk's 32 rows are 1 KB runs.  HBM takes such
            // rows at ~6.1 TB/s, 32-row x 128-byte tiles at ~5.2 (tools/dram_probe_rows.cu).  Every sample byte
            // becomes 4 bytes (util.inflate along x) and every sample row 4 pixel rows (along y); rows and columns
            // beyond the image are not written (the two crops). ----
            const int half = lane & 1;
            uint8_t* plane_ptr = a.planes_out + (size_t)plane * a.plane_stride;
            uint8_t* dstp[2];
            int rows_ok[2], cols_ok[2];
            #pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int b = 16 * h + (lane >> 1);
                const int blk = blk0 + b;
                const int by = blk / g.hb, bx = blk - by * g.hb;
                const int x0 = bx * 32 + 16 * half;
                dstp[h] = plane_ptr + (size_t)by * 32 * a.row_pitch + x0;
                rows_ok[h] = b < nvalid ? jb_min(32, g.H - by * 32) : 0;
                cols_ok[h] = jb_min(16, g.W - x0);
            }
            const uint8_t* sbase = wr.buf + (lane >> 1) * FI_SAMPLE_PITCH + 4 * half;
            #pragma unroll 2
            for (int m = 0; m < 8; ++m) {
                #pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t w = *(const uint32_t*)(sbase + 16 * h * FI_SAMPLE_PITCH + m * 8);
                    const uint4 v = make_uint4(__byte_perm(w, 0, 0x0000), __byte_perm(w, 0, 0x1111),
                                               __byte_perm(w, 0, 0x2222), __byte_perm(w, 0, 0x3333));
                    uint8_t* d = dstp[h] + (size_t)(4 * m) * a.row_pitch;
                    const int nrows = rows_ok[h] - 4 * m;
                    if (ka.aligned && cols_ok[h] == 16) {
                        #pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (k < nrows) *(uint4*)(d + (size_t)k * a.row_pitch) = v;
                    } else if (cols_ok[h] > 0) {
                        const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
                        for (int k = 0; k < 4 && k < nrows; ++k)
                            for (int q = 0; q < cols_ok[h]; ++q) d[(size_t)k * a.row_pitch + q] = (uint8_t)vv[q >> 2];
                    }
                }
            }
            __syncwarp();                                     // the sample buffer is rewritten by the next chunk
        }
        // ---- the pipeline moves on ----
        if (!next_located) prefetch_next();                             // (a chunk of fewer than four iterations)
        chunk = next_chunk;
        cur_k = next_k;
        pre_valid = next_pre;
    }
    if (!ROWS) {
        if (lane == 0) ff_bulk_wait_read<0>();
        __syncwarp();
    }
    jb_dec_epilogue(a);                                       // end of the call (jb_inverse.cuh)
}

bool jb_inv_fast_eligible(const JbGeom& g) { return g.d == 8 && g.bs == 4; }

template <bool DFT, int MODE, bool ROWS>
static cudaError_t jb_inv_fast_launch_t(const CUtensorMap& map, const FiKernelArgs& ka, cudaStream_t s) {
    const int NWARPS = ROWS ? FI_ROWS_WARPS : FI_WARPS;
    const size_t smem = 128 + (size_t)NWARPS * (ROWS ? sizeof(FiRowsSmem) : sizeof(FiWarpSmem));
    cudaError_t e = cudaFuncSetAttribute(jb_inv_fast_kernel<DFT, MODE, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jb_inv_fast_kernel<DFT, MODE, ROWS>, NWARPS * 32, smem);
    if (per_sm < 1) per_sm = 1;
    unsigned want = ka.a.n_chunks;
    unsigned grid = want < (unsigned)(sms * per_sm) ? want : (unsigned)(sms * per_sm);
    if (grid == 0) return cudaSuccess;
    return jb_launch_ex(jb_inv_fast_kernel<DFT, MODE, ROWS>, dim3(grid), dim3(NWARPS * 32), smem, s,
                        (ka.a.g.flags & JB_FLAG_PDL) != 0, map, ka);
}

template <bool ROWS>
static cudaError_t jb_inv_fast_launch_r(const CUtensorMap& map, const FiKernelArgs& ka, bool dft, int mode, cudaStream_t s) {
    if (mode == 0) return dft ? jb_inv_fast_launch_t<true, 0, ROWS>(map, ka, s) : jb_inv_fast_launch_t<false, 0, ROWS>(map, ka, s);
    return dft ? jb_inv_fast_launch_t<true, 2, ROWS>(map, ka, s) : jb_inv_fast_launch_t<false, 2, ROWS>(map, ka, s);
}

cudaError_t jb_launch_inv_fast(const JbInvArgs& a, int mode, cudaStream_t s) {
    FiKernelArgs ka;
    ka.a = a;
    const JbGeom& g = a.g;
    ka.aligned = (((uintptr_t)a.planes_out & 15) == 0 && (a.row_pitch & 15) == 0 &&
                  (a.n_planes == 1 || (a.plane_stride & 15) == 0)) ? 1 : 0;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    ka.use_tma = 0;
    const bool dft = g.transform == JB_TRANSFORM_DFT;
    // default: the chunk goes out row by row with 128-bit stores (no tile, no tensor map); JB_FLAG_TILE_DECODER keeps
    // the 4-block tiles, stored by TMA (or by plain stores with JB_FLAG_NO_TMA)
    if (!(g.flags & JB_FLAG_TILE_DECODER)) return jb_inv_fast_launch_r<true>(map, ka, dft, mode, s);
    if (!(g.flags & JB_FLAG_NO_TMA) && ka.aligned)
        ka.use_tma = jb_make_plane_tensor_map(&map, a.planes_out, g.W, g.H, a.n_planes, a.row_pitch, a.plane_stride) ? 1 : 0;
    return jb_inv_fast_launch_r<false>(map, ka, dft, mode, s);
}
