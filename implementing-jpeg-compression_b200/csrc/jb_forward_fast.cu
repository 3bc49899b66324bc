// jb_forward_fast.cu -- compress direction, specialised for dct_size 8 / block_size 4
// (the CLI default, BASELINE.json configs 1, 2, 4, 5), DCT or real-part DFT, any quantiser.
//
// Work unit: one warp owns a chunk of 32 consecutive blocks (raster order) of one plane.
//   * 8 iterations, each over a 32-row x 128-byte source tile = 4 horizontally adjacent
//     blocks.  Tiles are staged in shared memory by TMA (cp.async.bulk.tensor, one elected
//     lane, mbarrier completion) through a per-warp ring, so the next tiles are in flight
//     while the current one is transformed.  Tiles that touch the replicated edge
//     (padding.py:8-12, dct_padding.py:8-9) or a misaligned plane are filled by the warp
//     with clamped loads instead (TMA would zero-fill, the reference replicates).
//   * lane (i, b) = (row 0..7, block 0..3): 8 x LDS.128 -> 4x4 box sums by dp4a -> 8-point
//     transform along the row in registers -> transpose through shared memory -> 8-point
//     transform down the column -> quantise (fp32; coefficients that land within the
//     fp32 error bound of a rounding tie are re-evaluated in fp64 the way the reference
//     computes them) -> int16 store at the zigzag position.
//   * then lane t run-length-encodes and bit-packs block t of the chunk (util.py:146-160,
//     203-221) into a 64-byte staging row, the warp scans the 32 byte lengths, compacts the
//     rows in shared memory and writes the chunk with 128-bit stores to its slot of a
//     temporary buffer (blocks longer than 64 bytes are packed again, straight to the slot);
//     a device-wide exclusive scan of the chunk lengths and a gather (jb_forward.cu) then
//     place every chunk in the output stream.
// No block-level synchronisation: warps are independent after the table preload.
#include <cuda.h>
#include <string.h>

#include "jb_common.cuh"
#include "jb_fast_common.cuh"
#include "jb_forward.cuh"
#include "jb_refine.cuh"

#define FF_WARPS 16
#define FF_RING 2
#define FF_STAGE_W 17       // words per packed-bytes row: 64 bytes + 1 word (odd stride)
#define FF_STAGE_CAP 16     // words a block may occupy in its staging row; longer blocks take the slow path
#define FF_BIG_CAP 8
#define FF_COMPACT_BYTES (JB_CHUNK * FF_COEF_W * 4)   // the coefficient rows double as the compaction buffer

struct __align__(128) FfWarpSmem {
    uint8_t tile[FF_RING][FF_TILE_BYTES];   // the slot just consumed doubles as the packed-bytes staging rows
    float scr[4 * FF_BLK_W];            // row-pass results [block][row][col]; on demand also the box sums
    uint32_t coef[JB_CHUNK * FF_COEF_W];
    unsigned long long bar[FF_RING];
    int big_blk[FF_BIG_CAP], big_pos[FF_BIG_CAP], big_amp[FF_BIG_CAP];
    int nbig;
};

struct FfCtaSmem {
    double A64[64];
    double B64[64];
};

// ---- tile staging -----------------------------------------------------------------------------
// kind 0: the 32 x 128 tile lies inside the image, rows 16-byte aligned  -> TMA (LDG.128 when TMA is off)
// kind 1: same columns, but rows run past the bottom edge              -> LDG.128 with replicated rows
// kind 2: right edge, block-row wrap, partial group or unaligned plane -> clamped byte loads
struct FfCursor {           // position of a 4-block group inside the stream of chunks (warp-uniform)
    unsigned chunk;
    int plane, blk0, nvalid, nit, it;
    int by, bx;             // block coordinates of the group's first block
    int kind;               // ff_cursor_kind of this group (computed once, when the tile is issued)
};

__device__ __forceinline__ void ff_cursor_set(FfCursor& c, unsigned chunk, const JbGeom& g) {
    c.chunk = chunk;
    c.plane = (int)(chunk / (unsigned)g.cpp);
    c.blk0 = (int)(chunk - (unsigned)c.plane * (unsigned)g.cpp) * JB_CHUNK;
    c.nvalid = jb_min(JB_CHUNK, g.nblocks - c.blk0);
    c.nit = (c.nvalid + 3) >> 2;
    c.it = 0;
    c.by = c.blk0 / g.hb;
    c.bx = c.blk0 - c.by * g.hb;
}
__device__ __forceinline__ void ff_cursor_next(FfCursor& c, const JbGeom& g) {
    ++c.it;
    c.bx += 4;
    while (c.bx >= g.hb) { c.bx -= g.hb; ++c.by; }
}
__device__ __forceinline__ int ff_cursor_kind(const FfCursor& c, const JbGeom& g, bool aligned) {
    if (!aligned || 4 * c.it + 4 > c.nvalid || c.bx + 4 > g.hb || (c.bx + 4) * 32 > g.W) return 2;
    return (c.by + 1) * 32 <= g.H ? 0 : 1;
}

// two-level edge replication of SURVEY.md section 8(a) rows A1-A3, written into the tile layout
struct FfEdgeGeom { int hb, H, W, H1, W1; };     // by value: a reference to kernel parameters would force a local copy
__device__ __noinline__ void ff_fill_clamped(uint8_t* tile, const uint8_t* plane, size_t pitch, const FfEdgeGeom g,
                                             int n0, int nlive, int lane) {
    uint32_t* t32 = (uint32_t*)tile;
    for (int idx = lane; idx < 1024; idx += 32) {
        const int row = idx >> 5, wcol = idx & 31, b = wcol >> 3;
        uint32_t word = 0;
        if (b < nlive) {
            const int blk = n0 + b;
            const int by = blk / g.hb, bx = blk - by * g.hb;
            const int si = jb_min(by * 8 + (row >> 2), g.H1 - 1), sj = jb_min(bx * 8 + (wcol & 7), g.W1 - 1);
            const uint8_t* r = plane + (size_t)jb_min(si * 4 + (row & 3), g.H - 1) * pitch;
            const int x0 = sj * 4, xm = g.W - 1;
            word = (uint32_t)r[jb_min(x0, xm)] | ((uint32_t)r[jb_min(x0 + 1, xm)] << 8) |
                   ((uint32_t)r[jb_min(x0 + 2, xm)] << 16) | ((uint32_t)r[jb_min(x0 + 3, xm)] << 24);
        }
        t32[idx] = word;
    }
}

// whole columns inside the image: 128-bit loads, rows replicated past the bottom edge (kind 0 / 1)
__device__ __forceinline__ void ff_fill_rows(uint8_t* tile, const uint8_t* plane, size_t pitch, const JbGeom& g,
                                             int by, int bx, int lane) {
    const uint8_t* src = plane + (size_t)bx * 32;
    uint4* t16 = (uint4*)tile;
    uint4 v[8];
    #pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int idx = k * 32 + lane, row = idx >> 3;
        const int si = jb_min(by * 8 + (row >> 2), g.H1 - 1);
        const int y = jb_min(si * 4 + (row & 3), g.H - 1);
        v[k] = __ldg((const uint4*)(src + (size_t)y * pitch) + (idx & 7));
    }
    #pragma unroll
    for (int k = 0; k < 8; ++k) t16[k * 32 + lane] = v[k];
}

struct FfKernelArgs {
    JbFwdArgs a;
    int use_tma;        // tensor map valid
    int aligned;        // plane base / pitch / stride are multiples of 16 bytes
};

// Non-zero mask of one block's 64 int16 coefficients (zigzag order, row of FF_COEF_W words): per pair of 128-bit
// loads, min(halfword, 1) packed two to a word (VIMNMX.U16x2), the eight words added at even bit positions, and the
// high halfwords folded down between them.
__device__ __forceinline__ void ff_block_mask(const uint32_t* rowp, uint32_t& lo, uint32_t& hi) {
    uint32_t m16[4];
    #pragma unroll
    for (int q8 = 0; q8 < 4; ++q8) {
        const uint4 w0 = *(const uint4*)(rowp + 8 * q8), w1 = *(const uint4*)(rowp + 8 * q8 + 4);
        const uint32_t one = 0x00010001u;
        uint32_t m = __vminu2(w0.x, one) + (__vminu2(w0.y, one) << 2) + (__vminu2(w0.z, one) << 4) + (__vminu2(w0.w, one) << 6)
                   + (__vminu2(w1.x, one) << 8) + (__vminu2(w1.y, one) << 10) + (__vminu2(w1.z, one) << 12) + (__vminu2(w1.w, one) << 14);
        m16[q8] = (m | (m >> 15)) & 0xFFFFu;
    }
    lo = m16[0] | (m16[1] << 16);
    hi = m16[2] | (m16[3] << 16);
}

// Pack one block with any writer; amplitudes that do not fit their size field (BadRleCodeError in the reference) are
// skipped and the first of them reported.
template <typename Writer>
__device__ __forceinline__ void ff_pack_block(const uint32_t* rowp, Writer& bw, int& bad_pos, int& bad_run) {
    uint32_t lo, hi;
    ff_block_mask(rowp, lo, hi);
    const int16_t* c16 = (const int16_t*)rowp;
    int prev = -1, bad_prev = 0;
    bad_pos = -1;
    #pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        uint32_t m = half ? hi : lo;
        while (m) {
            const int p = half * 32 + __ffs((int)m) - 1;
            m &= m - 1;
            const int amp = c16[p];
            const int run = p - prev - 1;
            const uint32_t mag = (uint32_t)(amp < 0 ? -amp : amp);
            if (mag > (uint32_t)JB_MAX_AMP) {                    // BadRleCodeError in the reference: nothing is emitted
                if (bad_pos < 0) { bad_pos = p; bad_prev = prev; }
                prev = p;
                continue;
            }
            prev = p;
            int r = run;
            while (r >= JB_MAX_RUN) { bw.put(0xF0u, 8); r -= JB_MAX_RUN; }
            const int size = 33 - __clz((int)mag);               // bit length + 1 (mag >= 1)
            bw.put(((((uint32_t)r << 4) | (uint32_t)size) << size) | ((amp > 0 ? 1u : 0u) << (size - 1)) | mag, 8 + size);
        }
    }
    bad_run = bad_pos >= 0 ? (bad_pos - bad_prev - 1) % JB_MAX_RUN : 0;
    bw.put(0u, 8);
}

// The common case of the above -- every amplitude of the chunk fits (the quantiser stage counted none that does not) --
// into a staging row of FF_STAGE_W words: the run of the loop per non-zero coefficient is what the kernel spends a
// fifth of its instructions on, so it is kept branch-free.  Pending bits sit left-aligned in `acc`; every code stores
// the word it lands in (again, if the word is not full yet); words beyond FF_STAGE_CAP fall on the row's spare last
// word and are only counted.  Returns the block's length in bytes.
__device__ __forceinline__ unsigned ff_pack_block_fit(const uint32_t* rowp, uint32_t* out) {
    uint32_t lo, hi;
    ff_block_mask(rowp, lo, hi);
    const int16_t* c16 = (const int16_t*)rowp;
    uint32_t acc = 0;
    int nacc = 0, nw = 0;
    auto put = [&](uint32_t vl, int k) {                         // vl: the k <= 23 bits, left-aligned
        const uint32_t word = acc | (vl >> nacc);
        const uint32_t rest = __funnelshift_r(0u, vl, nacc);     // vl << (32 - nacc); 0 for nacc == 0
        out[jb_min(nw, FF_STAGE_CAP)] = jb_bswap32(word);
        const int t = nacc + k;
        acc = t >= 32 ? rest : word;
        nw += t >> 5;
        nacc = t & 31;
    };
    int off = 0;                                                 // run = off + b: minus the position behind the previous non-zero
    #pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        uint32_t m = half ? hi : lo;
        const int16_t* cb = c16 + 32 * half;
        while (m) {
            const int b = __ffs((int)m) - 1;
            m &= m - 1;
            const int amp = cb[b];
            int run = off + b;
            off = ~b;                                            // -(b + 1)
            #pragma unroll 1
            while (run >= JB_MAX_RUN) { put(0xF0000000u, 8); run -= JB_MAX_RUN; }     // (rare: at most three times)
            const uint32_t mag = (uint32_t)(amp < 0 ? -amp : amp);
            const int sz1 = 32 - __clz((int)mag);                // bit length = size - 1
            const uint32_t head = ((uint32_t)run << 5) | ((uint32_t)(sz1 + 1) << 1) | (amp > 0 ? 1u : 0u);
            put((head << 23) | (mag << (23 - sz1)), 9 + sz1);
        }
        off += 32;
    }
    put(0u, 8);
    out[jb_min(nw, FF_STAGE_CAP)] = jb_bswap32(acc);
    return ((unsigned)nw * 32u + (unsigned)nacc + 7u) >> 3;
}

// ---- the kernel -------------------------------------------------------------------------------
template <bool DFT, int MODE>
__global__ void __launch_bounds__(FF_WARPS * 32, 1)
jb_fwd_fast_kernel(const __grid_constant__ CUtensorMap tmap, const FfKernelArgs ka) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const JbFwdArgs& a = ka.a;
    const JbGeom& g = a.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    FfCtaSmem& cs = *(FfCtaSmem*)smem_raw;
    FfWarpSmem& ws = *(FfWarpSmem*)(smem_raw + 1024 + (size_t)warp * sizeof(FfWarpSmem));
    jb_pdl_trigger();
    // the tables are constant once built (JB_FLAG_REUSE_TABLES): their loads may overlap the kernel before this one
    const bool tables_const = (g.flags & JB_FLAG_REUSE_TABLES) != 0;
    if (!tables_const) jb_pdl_wait();

    for (int i = threadIdx.x; i < 64; i += blockDim.x) { cs.A64[i] = a.t.fA64[i]; cs.B64[i] = a.t.fB64[i]; }
    if (lane == 0) {
        for (int s = 0; s < FF_RING; ++s) ff_mbar_init(&ws.bar[s], 1);
        ws.nbig = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // lane roles: (i, b) for loads and the row pass, (c, b) for the column pass
    const int li = lane >> 2, lb = lane & 3;
    const int v = DFT ? (li < 5 ? li : 12 - li) : li;          // frequency column this lane quantises
    float qm[8];
    uint32_t zzlo = 0, zzhi = 0;                               // zigzag positions of (u, v), one byte each
    #pragma unroll
    for (int u = 0; u < 8; ++u) {
        qm[u] = a.t.qmult[u * 8 + v];
        const uint32_t z = a.t.zz[u * 8 + v];
        if (u < 4) zzlo |= z << (8 * u); else zzhi |= z << (8 * (u - 4));
    }
    // near a tie  <=>  |frac - .5| < tol + 2.4e-7 |val|, tol = qtol[u][v] the bound jb_tables.cu derived for the fp32 path:
    // thr[u] = .5 - tol, compared with |val - rint(val)| + 2.4e-7 |val|
    float thr[8];
    #pragma unroll
    for (int u = 0; u < 8; ++u) {
        const float m0 = fabsf(a.t.qmult[u * 8 + v]);
        const float tol = m0 > 0.f ? (a.t.qtol[u * 8 + v] - 1e-7f) / m0 * m0 : 0.f;
        thr[u] = 0.5f - 1e-7f - tol;
    }
    const bool refine_on = !(g.flags & JB_FLAG_NO_REFINE);
    const bool aligned = ka.aligned != 0;
    const bool use_tma = ka.use_tma != 0;

    if (tables_const) jb_pdl_wait();                           // from here on: data the kernels before this one wrote
    const int P = jb_ctrl_parity(a);                           // which set of the control block this call uses
    unsigned* const ticket = jb_ctrl_ticket(a, P);
    // The first chunk of every warp is dealt statically, CTA-major (warp w of CTA b: chunk b + gridDim w): a small input
    // (one 4K frame: 765 chunks) then spreads over all SMs instead of filling the first few; the chunks behind that first
    // round are claimed from the ticket counter.
    const unsigned static_chunks = gridDim.x * FF_WARPS;
    if (blockIdx.x == 0 && threadIdx.x == 0) a.ctrl[1] = (unsigned)P;        // the parity, for the kernels behind this one
    auto claim = [&]() -> unsigned {
        unsigned c = 0;
        if (lane == 0) c = static_chunks + atomicAdd(ticket, 1u);
        return __shfl_sync(0xffffffffu, c, 0);
    };

    // The warp walks a stream of 4-block tiles (chunk after chunk) one tile ahead of itself: while tile
    // (bx, by) of chunk `ck` is transformed, the TMA load of the next tile is in flight in the other ring
    // slot -- the next tile of the same chunk, or the first tile of a freshly claimed chunk `nk`.
    // Chunk-level state only moves at chunk boundaries; per tile only (bx, by, kind) do.
    FfCursor ck, nk;
    {
        const unsigned first = blockIdx.x + gridDim.x * (unsigned)warp;
        if (first >= a.n_chunks) return;
        ff_cursor_set(ck, first, g);
    }
    const int hb = g.hb;
    const int bx_lim = aligned ? jb_min(hb - 4, g.W / 32 - 4) : -1;      // beyond it: wrap / right edge / unaligned -> kind 2
    const int by_lim = g.H / 32 - 1;                                     // beyond it: rows replicated -> kind 1
    auto tile_kind = [&](const FfCursor& c, int it, int bx, int by) -> int {
        if (bx > bx_lim || 4 * it + 4 > c.nvalid) return 2;
        return by <= by_lim ? 0 : 1;
    };
    // Generic-proxy writes into a ring slot (edge-tile fills, the staging rows of the pack stage) must be
    // ordered before the TMA engine writes the slot again; tiles that were only read need no proxy fence.
    bool dirty = true;
    auto issue = [&](int plane, int bx, int by, int kind, int slot) {
        if (use_tma && kind == 0 && lane == 0) {
            if (dirty) ff_fence_proxy_async();
            ff_mbar_expect_tx(&ws.bar[slot], FF_TILE_BYTES);
            ff_tma_load_3d(ws.tile[slot], &tmap, bx * 32, by * 32, plane, &ws.bar[slot]);
        }
    };
    int slot = 0;
    unsigned phasebits = 0;
    int bx = ck.bx, by = ck.by;
    int kind = tile_kind(ck, 0, bx, by);
    issue(ck.plane, bx, by, kind, 0);
    bool have = true;

    while (have) {
      bool have_next = true;
      #pragma unroll 1
      for (int it = 0; it < ck.nit; ++it) {
        // ---- the tile after this one: same chunk, or the first tile of a freshly claimed chunk ----
        int nbx, nby, nkind, nplane = ck.plane;
        if (it + 1 < ck.nit) {
            nbx = bx + 4; nby = by;
            while (nbx >= hb) { nbx -= hb; ++nby; }
            nkind = tile_kind(ck, it + 1, nbx, nby);
        } else {
            const unsigned c2 = claim();
            have_next = c2 < a.n_chunks;
            nbx = nby = 0; nkind = 2;
            if (have_next) {
                ff_cursor_set(nk, c2, g);
                nbx = nk.bx; nby = nk.by; nplane = nk.plane;
                nkind = tile_kind(nk, 0, nbx, nby);
            }
        }
        if (have_next) issue(nplane, nbx, nby, nkind, slot ^ 1);
        dirty = false;

        {
            uint8_t* tile = ws.tile[slot];
            if (kind == 0 && use_tma) {
                const uint32_t par = (phasebits >> slot) & 1u;
                while (!ff_mbar_try_wait(&ws.bar[slot], par)) { }
                phasebits ^= 1u << slot;
            } else if (kind <= 1) {
                const uint8_t* plane_ptr = a.planes + (size_t)ck.plane * a.plane_stride;
                ff_fill_rows(tile, plane_ptr, a.row_pitch, g, by, bx, lane);
                dirty = true;
                __syncwarp();
            } else {
                const uint8_t* plane_ptr = a.planes + (size_t)ck.plane * a.plane_stride;
                const FfEdgeGeom eg = {g.hb, g.H, g.W, g.H1, g.W1};
                ff_fill_clamped(tile, plane_ptr, a.row_pitch, eg, ck.blk0 + 4 * it, jb_min(4, ck.nvalid - 4 * it), lane);
                dirty = true;
                __syncwarp();
            }

            // ---- box sums: rows 4i..4i+3, bytes 32b..32b+31 of the tile ----
            float x[8];
            {
                const uint4* t16 = (const uint4*)tile;
                const int h0 = li & 1;                      // alternate the 16-byte half: conflict-free LDS.128
                uint32_t sa[4] = {0x4B000000u, 0x4B000000u, 0x4B000000u, 0x4B000000u};
                uint32_t sb[4] = {0x4B000000u, 0x4B000000u, 0x4B000000u, 0x4B000000u};
                #pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int row = 4 * li + k;
                    const uint4 qa = t16[row * 8 + 2 * lb + h0];
                    const uint4 qb = t16[row * 8 + 2 * lb + (h0 ^ 1)];
                    sa[0] = __dp4a(qa.x, 0x01010101u, sa[0]); sa[1] = __dp4a(qa.y, 0x01010101u, sa[1]);
                    sa[2] = __dp4a(qa.z, 0x01010101u, sa[2]); sa[3] = __dp4a(qa.w, 0x01010101u, sa[3]);
                    sb[0] = __dp4a(qb.x, 0x01010101u, sb[0]); sb[1] = __dp4a(qb.y, 0x01010101u, sb[1]);
                    sb[2] = __dp4a(qb.z, 0x01010101u, sb[2]); sb[3] = __dp4a(qb.w, 0x01010101u, sb[3]);
                }
                // 0x4B000000 + s is the bit pattern of 2^23 + s: exact int -> float without I2F
                #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float fa = __uint_as_float(sa[j]) - 8388608.0f, fb = __uint_as_float(sb[j]) - 8388608.0f;
                    x[j] = h0 ? fb : fa;
                    x[4 + j] = h0 ? fa : fb;
                }
            }
            // ---- row pass, transpose through shared memory, column pass ----
            float r[8];
            if (DFT) ff_rdft8(x, r); else ff_dct8(x, r);
            const int h0 = li & 1;
            float4* sc = (float4*)(ws.scr + lb * FF_BLK_W + li * 8);
            sc[h0] = h0 ? make_float4(r[4], r[5], r[6], r[7]) : make_float4(r[0], r[1], r[2], r[3]);
            sc[h0 ^ 1] = h0 ? make_float4(r[0], r[1], r[2], r[3]) : make_float4(r[4], r[5], r[6], r[7]);
            __syncwarp();
            float col[8], y[8];
            #pragma unroll
            for (int k = 0; k < 8; ++k) col[k] = ws.scr[lb * FF_BLK_W + k * 8 + li];
            if (DFT) ff_dft_column_stage(col, li, y); else ff_dct8(col, y);

            // ---- quantise, tie check ----
            const int gblk = 4 * it + lb;                // block inside the chunk
            int qi[8];
            unsigned nearmask = 0;
            float vmax = 0.f;
            #pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float val = y[u] * qm[u];
                vmax = fmaxf(vmax, fabsf(val));
                const float t = val + 12582912.0f;          // 1.5 * 2^23: rounds half-even to an integer
                const float dd = val - (t - 12582912.0f);
                qi[u] = __float_as_int(t) - 0x4B400000;
                const float e = fmaf(fabsf(val), 2.4e-7f, fabsf(dd));
                if (e > thr[u]) nearmask |= 1u << u;
            }
            if (refine_on && __any_sync(0xffffffffu, nearmask != 0)) {
                // the box sums go back to shared memory only now: nearly every iteration skips this
                __syncwarp();
                sc[h0] = h0 ? make_float4(x[4], x[5], x[6], x[7]) : make_float4(x[0], x[1], x[2], x[3]);
                sc[h0 ^ 1] = h0 ? make_float4(x[0], x[1], x[2], x[3]) : make_float4(x[4], x[5], x[6], x[7]);
                __syncwarp();
                // Where every twiddle is 0 or +-1 -- the DC term of the DCT; rows and columns 0, 2, 4, 6 of the DFT -- the
                // reference's float64 transform is an exact sum of the box means, and so is y[u] (integers below 2^24):
                // only the quantiser step has to be redone in float64.  These are the positions where exact ties are
                // common (K/256 for the DC term).  The other positions: one coefficient at a time
                // by the whole warp (jb_pf_fft2_8x8_warp / jb_dct2_8x8_warp), the lanes taking turns -- a lane on its own
                // would keep the other 31 waiting five to ten times as long (5 % of the DCT kernel's instructions).
                unsigned pend = 0;
                #pragma unroll
                for (int u = 0; u < 8; ++u)
                    if ((nearmask >> u & 1u) && (DFT ? (((u | v) & 1) != 0) : ((u | v) != 0))) pend |= 1u << u;
                nearmask &= ~pend;
                for (;;) {
                    const unsigned have = __ballot_sync(0xffffffffu, pend != 0);
                    if (!have) break;
                    const int L = __ffs((int)have) - 1;
                    const int u = __ffs((int)__shfl_sync(0xffffffffu, pend, L)) - 1;
                    const int vL = __shfl_sync(0xffffffffu, v, L);
                    const float* XL = ws.scr + (L & 3) * FF_BLK_W;
                    const double F = DFT ? jb_pf_fft2_8x8_warp(XL, u, vL) : jb_dct2_8x8_warp(XL, u, vL, cs.A64);
                    if (lane == L) {
                        const double recip = a.t.qrecip[u * 8 + v];
                        const double f64 = g.qmode == JB_Q_QTABLE ? __dmul_rn(F, recip) : g.qmode == JB_Q_DIVIDE ? __ddiv_rn(F, recip) : F;
                        const int q = (int)rint(f64);
                        #pragma unroll
                        for (int uu = 0; uu < 8; ++uu) if (uu == u) qi[uu] = q;
                        pend &= pend - 1u;
                    }
                }
                if (nearmask) {
                    #pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (nearmask >> u & 1u) {
                            const double recip = a.t.qrecip[u * 8 + v];
                            const double F = __dmul_rn((double)y[u], 0.0625);
                            const double f64 = g.qmode == JB_Q_QTABLE ? __dmul_rn(F, recip) : g.qmode == JB_Q_DIVIDE ? __ddiv_rn(F, recip) : F;
                            qi[u] = (int)rint(f64);
                        }
                }
                __syncwarp();
            }
            // ---- zigzag store ----
            if (gblk < ck.nvalid) {
                if (MODE == 1) {
                    int16_t* dst = a.coeffs_out + ((size_t)ck.plane * g.nblocks + ck.blk0 + gblk) * 64;
                    #pragma unroll
                    for (int u = 0; u < 8; ++u)
                        dst[((u < 4 ? zzlo : zzhi) >> (8 * (u & 3))) & 0xFFu] = (int16_t)max(-32767, min(32767, qi[u]));
                } else {
                    int16_t* row = (int16_t*)(ws.coef + gblk * FF_COEF_W);
                    if (vmax > (float)JB_MAX_AMP - 0.75f) {
                        // some amplitude may not fit the 15-bit size field: remember the true values for
                        // the error report, store saturated ones (run_length_encoding stage will flag them)
                        #pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int zp = ((u < 4 ? zzlo : zzhi) >> (8 * (u & 3))) & 0xFFu;
                            int q = qi[u];
                            if (q > JB_MAX_AMP || q < -JB_MAX_AMP) {
                                const int k = atomicAdd(&ws.nbig, 1);
                                if (k < FF_BIG_CAP) { ws.big_blk[k] = gblk; ws.big_pos[k] = zp; ws.big_amp[k] = q; }
                                qi[u] = q > 0 ? 32767 : -32767;
                            }
                        }
                    }
                    #pragma unroll
                    for (int u = 0; u < 8; ++u)
                        row[((u < 4 ? zzlo : zzhi) >> (8 * (u & 3))) & 0xFFu] = (int16_t)qi[u];
                }
            }
            __syncwarp();          // tile and scr are free again
        }

        if (MODE == 0 && it + 1 == ck.nit) {
            // ---- A9 + A10: lane t packs block t into its (small) staging row ----
            // (the chunk's last tile has been consumed: its ring slot holds the staging rows until the next
            // TMA load is issued into it, behind a proxy fence)
            uint32_t* stage = (uint32_t*)ws.tile[slot];
            dirty = true;
            unsigned len = 0;
            if (ws.nbig == 0) {
                // (no amplitude of this chunk is too large for its size field: the lean packer)
                if (lane < ck.nvalid) len = ff_pack_block_fit(ws.coef + lane * FF_COEF_W, stage + lane * FF_STAGE_W);
            } else if (lane < ck.nvalid) {
                JbBitWriter bw;
                bw.init(stage + lane * FF_STAGE_W, FF_STAGE_CAP);
                int bad_pos, bad_run;
                ff_pack_block(ws.coef + lane * FF_COEF_W, bw, bad_pos, bad_run);
                len = bw.finish();
                if (bad_pos >= 0) {
                    long long amp = ((const int16_t*)(ws.coef + lane * FF_COEF_W))[bad_pos];
                    const int nb = jb_min(ws.nbig, FF_BIG_CAP);
                    for (int k = 0; k < nb; ++k)
                        if (ws.big_blk[k] == lane && ws.big_pos[k] == bad_pos) amp = ws.big_amp[k];
                    jb_report_bad_code(jb_ctrl_status(a, P), (unsigned long long)ck.plane * g.nblocks + ck.blk0 + lane,
                                       bad_pos, bad_run, amp);
                }
            }
            unsigned incl = len;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
            const unsigned excl = incl - len;
            if (lane == 0) jb_record_chunk_len(a, P, ck.chunk, total);
            uint8_t* slot = jb_chunk_slot(a, ck.chunk, total);
            const bool small = total <= FF_COMPACT_BYTES && !__any_sync(0xffffffffu, len > FF_STAGE_CAP * 4);
            if (small) {
                // compact the 32 rows inside shared memory (the coefficient rows are dead now), then
                // write the chunk with 128-bit stores; the slot is 16-byte aligned
                __syncwarp();
                uint8_t* cbuf = (uint8_t*)ws.coef;
                // lane t moves its row to byte `excl`: single bytes up to the first word boundary of the destination,
                // then whole words funnelled out of pairs of source words, then the last bytes (the neighbours' rows
                // share the first and the last word)
                const uint32_t* sw = stage + lane * FF_STAGE_W;
                const uint8_t* sb = (const uint8_t*)sw;
                unsigned j = 0, d = excl;
                while ((d & 3u) && j < len) cbuf[d++] = sb[j++];
                if (j < len) {
                    const unsigned nfull = (len - j) >> 2, sh = 8u * j;          // (j < 4 here)
                    uint32_t* dw = (uint32_t*)(cbuf + d);
                    uint32_t cur = sw[0];
                    for (unsigned k = 0; k < nfull; ++k) {
                        const uint32_t nxt = sw[k + 1];                          // (at most word FF_STAGE_CAP of the row)
                        dw[k] = __funnelshift_r(cur, nxt, sh);
                        cur = nxt;
                    }
                    j += 4u * nfull; d += 4u * nfull;
                    while (j < len) cbuf[d++] = sb[j++];
                }
                __syncwarp();
                const uint4* c16 = (const uint4*)cbuf;
                for (unsigned i = lane; i < (total + 15u) >> 4; i += 32) ((uint4*)slot)[i] = c16[i];
            } else {
                // a block longer than its staging row (or a very dense chunk): pack again, bytes straight
                // to the slot
                if (lane < ck.nvalid) {
                    JbByteWriter bw;
                    bw.init(slot + excl);
                    int bad_pos, bad_run;
                    ff_pack_block(ws.coef + lane * FF_COEF_W, bw, bad_pos, bad_run);
                    bw.finish();
                }
            }
            if (lane == 0) ws.nbig = 0;
            __syncwarp();
        }

        bx = nbx; by = nby; kind = nkind;
        slot ^= 1;
      }
      have = have_next;
      ck = nk;
    }
}

// ---- host side ----------------------------------------------------------------------------------
bool jb_fwd_fast_eligible(const JbGeom& g) { return g.d == 8 && g.bs == 4; }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled jb_get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// 3-D uint8 tensor (x = byte in row, y = row, z = plane), box 128 x 32 x 1
bool jb_make_plane_tensor_map(CUtensorMap* map, const void* base, int W, int H, int n_planes, size_t row_pitch,
                              size_t plane_stride, int box_w, int box_h) {
    PFN_encodeTiled enc = jb_get_encode_tiled();
    if (!enc) return false;
    if (((uintptr_t)base & 15) || (row_pitch & 15) || (plane_stride & 15) || W < box_w || H < box_h) return false;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_planes};
    cuuint64_t strides[2] = {(cuuint64_t)row_pitch, (cuuint64_t)(n_planes > 1 ? plane_stride : row_pitch * (size_t)H)};
    if (strides[1] & 15) return false;
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <bool DFT, int MODE>
static cudaError_t jb_fwd_fast_launch_t(const CUtensorMap& map, const FfKernelArgs& ka, cudaStream_t s) {
    const size_t smem = 1024 + (size_t)FF_WARPS * sizeof(FfWarpSmem);
    cudaError_t e = cudaFuncSetAttribute(jb_fwd_fast_kernel<DFT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jb_fwd_fast_kernel<DFT, MODE>, FF_WARPS * 32, smem);
    if (per_sm < 1) per_sm = 1;
    unsigned want = ka.a.n_chunks;                              // (warps without a chunk leave at once)
    unsigned grid = want < (unsigned)(sms * per_sm) ? want : (unsigned)(sms * per_sm);
    if (grid == 0) return cudaSuccess;
    return jb_launch_ex(jb_fwd_fast_kernel<DFT, MODE>, dim3(grid), dim3(FF_WARPS * 32), smem, s,
                        (ka.a.g.flags & JB_FLAG_PDL) != 0, map, ka);
}

cudaError_t jb_launch_fwd_fast(const JbFwdArgs& a, int mode, cudaStream_t s) {
    FfKernelArgs ka;
    ka.a = a;
    const JbGeom& g = a.g;
    ka.aligned = (((uintptr_t)a.planes & 15) == 0 && (a.row_pitch & 15) == 0 &&
                  (a.n_planes == 1 || (a.plane_stride & 15) == 0)) ? 1 : 0;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    ka.use_tma = 0;
    if (!(g.flags & JB_FLAG_NO_TMA) && ka.aligned)
        ka.use_tma = jb_make_plane_tensor_map(&map, a.planes, g.W, g.H, a.n_planes, a.row_pitch, a.plane_stride) ? 1 : 0;
    const bool dft = g.transform == JB_TRANSFORM_DFT;
    if (mode == 0) {
        cudaError_t e = dft ? jb_fwd_fast_launch_t<true, 0>(map, ka, s) : jb_fwd_fast_launch_t<false, 0>(map, ka, s);
        if (e != cudaSuccess) return e;
        return jb_launch_gather(a, s);
    }
    return dft ? jb_fwd_fast_launch_t<true, 1>(map, ka, s) : jb_fwd_fast_launch_t<false, 1>(map, ka, s);
}
