// jb_framing.cu -- parallel discovery of block boundaries in RleBytestream output.
//
// The container stores one length per band and nothing else (file_format.py:86-93); a block
// ends wherever its EOB lands plus zero padding to the byte (rle_byte_stream.py:55-56,80-82),
// so the reference finds block k only by decoding blocks 0..k-1 (BitDecoder,
// rle_byte_stream.py:6-42).  Here:
//
//   * every block ends with a 0x00 byte (the EOB's eight zero bits reach the end of a byte and
//     the padding is zero), so a block can only start at offset 0 or right after a 0x00 byte;
//   * each stream is cut into tiles of T bytes (256 for typical content).  WALK: one thread per tile starts
//     at the first such offset in its tile and follows block extents to the tile end, recording every start
//     it visits (one bit per byte of the tile) and where it leaves the tile (its exit; a long block may carry it over several tiles).
//     A warp owns 32 consecutive tiles of one stream and stages them in shared memory as one contiguous region;
//     the walk reads two codes per pass from one 32-bit window.
//     The walk of tile 0 starts at offset 0, which is a true block start; a walk that starts on a
//     false offset re-synchronises with the true chain within a block or two (it lands after a 0x00
//     byte, and nearly all of those are true ends);
//   * REACH: the tiles in which a true block starts form a chain: tile 0, then the tile holding the
//     exit of tile 0's walk, and so on.  Per stream that chain is marked by pointer doubling over the
//     "tile of my exit" links, and every tile on it receives its entry E = the exit of its predecessor;
//   * LINK: one thread per tile on the chain checks E against its own walk: if E is among the starts
//     the walk visited, everything from there on is true; if not, the thread follows the chain from
//     E until it meets the walk (a few blocks).  Either way the tile's true exit equals its walk's
//     exit, so by induction from tile 0 every entry is true -- and all tiles are linked in parallel.
//     A tile whose entry never meets the walk marks its stream for the serial fallback;
//   * SCAN: per stream, an exclusive scan of the per-tile block counts gives each tile the
//     ordinal of its first block; EMIT writes block_start[ordinal] = offset;
//   * SERIAL: a stream that failed any check is walked once by a single thread (correct for
//     every valid stream, just slow) and reports JB_ERR_BAD_STREAM if that fails too.
#include "jb_common.cuh"
#include "jb_inverse.cuh"

#define JB_POS_INVALID 0xFFFFFFFFu

// ---- sequential extent parser: aligned 32-bit big-endian-assembled words, 64-bit window ---------
// The word source is either global memory (any tile size) or a per-thread slice of shared memory
// that the warp filled with coalesced loads (small tiles: keeps thousands of concurrent walks
// from thrashing L1 with one cache line each).
struct JbWalker {
    const uint32_t* words;      // 4-byte aligned base at or below byte `origin` of the stream
    uint32_t origin;            // stream byte offset that words[0] (after shift) corresponds to
    uint32_t shift_bits;        // bit offset of byte `origin` inside words[]
    uint32_t nwords;            // words that may be read
    uint32_t len_bits;          // stream length in bits
    uint32_t widx;
    uint64_t buf;
    int nb;
    uint32_t bp;                // bit position inside the stream of the next unread bit
    bool in_smem;
    const uint32_t* gwords;     // staged window only: the same words in global memory ...
    uint32_t gnwords;           // ... which reach further (a block longer than the staged halo)

    __device__ __forceinline__ uint32_t load(uint32_t i) const {
        if (in_smem) {
            if (i < nwords) return jb_bswap32(words[i]);
            return i < gnwords ? jb_bswap32(__ldg(gwords + i)) : 0u;
        }
        return i < nwords ? jb_bswap32(__ldg(words + i)) : 0u;
    }
    // whole stream in global memory
    __device__ __forceinline__ void init(const uint8_t* stream, uint32_t len_bytes) {
        const uintptr_t a = (uintptr_t)stream;
        words = (const uint32_t*)(a & ~(uintptr_t)3);
        origin = 0;
        shift_bits = (uint32_t)(a & 3) * 8u;
        nwords = (uint32_t)((shift_bits / 8u + len_bytes + 3u) >> 2);
        len_bits = len_bytes * 8u;
        in_smem = false;
    }
    // a staged window: smem_words[0] holds the aligned word that contains stream byte `first_byte`
    __device__ __forceinline__ void init_window(const uint32_t* smem_words, uint32_t n_words,
                                                const uint32_t* global_words, uint32_t n_global_words,
                                                uint32_t first_byte, uint32_t misalign_bytes, uint32_t len_bytes) {
        words = smem_words;
        gwords = global_words;
        gnwords = n_global_words;
        origin = first_byte;
        shift_bits = misalign_bytes * 8u;
        nwords = n_words;
        len_bits = len_bytes * 8u;
        in_smem = true;
    }
    __device__ __forceinline__ uint32_t byte_at(uint32_t pos) const {       // stream byte `pos` (>= origin)
        const uint32_t b = pos - origin + (shift_bits >> 3);
        const uint32_t w = (in_smem && (b >> 2) < nwords) ? words[b >> 2] : __ldg((in_smem ? gwords : words) + (b >> 2));
        return (w >> ((b & 3u) * 8u)) & 0xFFu;
    }
    __device__ __forceinline__ void seek(uint32_t byte_pos) {
        bp = byte_pos * 8u;
        const uint32_t abs_bits = (byte_pos - origin) * 8u + shift_bits;
        widx = abs_bits >> 5;
        buf = load(widx++);
        nb = 32 - (int)(abs_bits & 31u);
    }
    // Parse one block that starts at the current (byte-aligned) position.  Returns false if it
    // is malformed; on success the reader stands at the first byte after the block.
    __device__ __forceinline__ bool block(int n, uint32_t maxblk_bits) {
        const uint32_t start = bp;
        int count = 0;
        for (;;) {                      // every pass adds at least 1 to `count`: at most n + 1 passes
            if (nb < 23) { buf = (buf << 32) | load(widx++); nb += 32; }
            const uint32_t head = (uint32_t)(buf >> (nb - 8)) & 0xFFu;
            const uint32_t size = head & 15u;
            if (size == 0u) {
                nb -= 8;
                if (head == 0u) break;                                   // EOB
                if (head != 0xF0u) return false;                         // (r, 0), 0 < r < 15
                count += JB_MAX_RUN;
            } else {
                if (size == 1u) return false;
                count += (int)(head >> 4) + 1;
                nb -= 8 + (int)size;
            }
            if (count > n) return false;
        }
        // bits consumed from words[0] = 32 * widx - nb; skip the zero padding to the byte boundary
        uint32_t abs_bits = widx * 32u - (uint32_t)nb;
        const uint32_t pad = (8u - (abs_bits & 7u)) & 7u;
        nb -= (int)pad;
        abs_bits += pad;
        bp = origin * 8u + abs_bits - shift_bits;
        return bp <= len_bits && bp - start <= maxblk_bits;
    }
};

__device__ __forceinline__ int jb_stream_of_tile(const unsigned* tile_first, int n_planes, unsigned tile) {
    int lo = 0, hi = n_planes;                 // largest s < n_planes with tile_first[s] <= tile
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (tile_first[mid] <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- F0: tiles per stream ----------------------------------------------------------------------
#define JB_REACH_SMALL_CAP 4096u
__global__ void __launch_bounds__(1024) jb_frame_prep_kernel(JbFrameArgs f) {
    __shared__ unsigned s_warp[33];
    __shared__ unsigned s_big[32][3];
    __shared__ int s_nbig;
    unsigned carry = 0, wcarry = 0;
    bool bad = false;
    if (threadIdx.x == 0) { f.big_list[0] = 0u; s_nbig = 0; }
    // (status words and chunk ticket live in the workspace's control block, which is clean when a call starts)
    __syncthreads();
    for (int base = 0; base < f.n_planes; base += 1024) {
        int s = base + threadIdx.x;
        unsigned nt = 0;
        if (s < f.n_planes) {
            unsigned long long len = f.plane_len[s];
            if (len > 0x1FFFFFF0ull) { bad = true; len = 0; }           // bit positions are 32-bit
            nt = (unsigned)((len + f.tile_bytes - 1) / f.tile_bytes);
            if (len > f.in_bytes || f.plane_off[s] > f.in_bytes - len) { bad = true; nt = 0; }     // reaches beyond the input buffer
            f.fallback[s] = f.force_serial ? 1u : 0u;
            f.notplain[s] = 0u;
            if (nt > JB_REACH_SMALL_CAP) f.big_list[1 + atomicAdd(f.big_list, 1u)] = (unsigned)s;
        }
        unsigned total;
        unsigned ex = jb_block_excl_scan(nt, s_warp, &total);
        if (s < f.n_planes) f.tile_first[s] = carry + ex;
        carry += total;
        // the walk gives every stream whole warps (32 tiles each)
        ex = jb_block_excl_scan((nt + 31u) >> 5, s_warp, &total);
        if (s < f.n_planes) {
            f.warp_first[s] = wcarry + ex;
            const unsigned wmax = f.max_tiles / 32u + (unsigned)f.n_planes + 1u;            // capacity of warp_stream
            const unsigned nw = (nt + 31u) >> 5;
            int slot = -1;
            if (nw > 64u) slot = atomicAdd(&s_nbig, 1);          // a long stream: the whole CTA fills its range below
            if (slot >= 0 && slot < 32) {
                s_big[slot][0] = (unsigned)s; s_big[slot][1] = wcarry + ex; s_big[slot][2] = nw;
            } else {
                for (unsigned j = 0, w = wcarry + ex; j < nw && w < wmax; ++j, ++w) f.warp_stream[w] = (unsigned)s;
            }
        }
        wcarry += total;
        __syncthreads();
        {
            const unsigned wmax = f.max_tiles / 32u + (unsigned)f.n_planes + 1u;
            const int nbig = s_nbig < 32 ? s_nbig : 32;
            for (int b = 0; b < nbig; ++b)
                for (unsigned j = threadIdx.x; j < s_big[b][2] && s_big[b][1] + j < wmax; j += blockDim.x)
                    f.warp_stream[s_big[b][1] + j] = s_big[b][0];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_nbig = 0;
        __syncthreads();
    }
    if (__syncthreads_or(bad ? 1 : 0)) carry = 0xFFFFFFFFu;
    if (threadIdx.x == 0) {
        if (carry > f.max_tiles) {
            jb_set_error(f.status, JB_ERR_BAD_STREAM);
            carry = 0;
        }
        f.tile_first[f.n_planes] = carry;
        f.warp_first[f.n_planes] = carry ? wcarry : 0u;
    }
}

// ---- F1: walk ----------------------------------------------------------------------------------
#define JB_WALK_THREADS 128
#define JB_WALK_HALO 64           // bytes staged beyond the tile; longer blocks continue from global memory

// The usual stream: every walk leaves its tile for the next one (it takes a block longer than a whole tile to skip
// one), i.e. every tile is on the chain of tiles and its entry is the exit of the tile before it.  The walks record
// exactly that -- entry of tile t + 1 = exit of tile t -- and raise the stream's flag where it does not hold; only then
// does jb_reach_stream work the chain out (and rewrite every entry).  t, nt: tile index and count within the stream.
__device__ __forceinline__ void jb_walk_note_exit(const JbFrameArgs& f, int s, unsigned tile, unsigned t, unsigned nt,
                                                  uint32_t exit_pos, uint32_t len) {
    const unsigned tsh = (unsigned)__ffs((int)f.tile_bytes) - 1u;
    bool ok;
    if (t + 1u < nt) {
        ok = exit_pos != JB_POS_INVALID && exit_pos < len && (exit_pos >> tsh) == t + 1u;
        if (ok) f.tile_entry[tile + 1u] = exit_pos;
    } else {
        ok = exit_pos == len;
    }
    if (t == 0u) f.tile_entry[tile] = 0u;
    if (!ok) f.notplain[s] = 1u;
}

// any tile size: every thread reads its tile straight from global memory and keeps its bitmap there
__global__ void __launch_bounds__(JB_WALK_THREADS) jb_frame_walk_kernel(JbFrameArgs f) {
    const unsigned tile = blockIdx.x * JB_WALK_THREADS + threadIdx.x;
    if (tile >= f.tile_first[f.n_planes]) return;
    const int s = jb_stream_of_tile(f.tile_first, f.n_planes, tile);
    const uint32_t len = (uint32_t)f.plane_len[s];
    const uint32_t tstart = (tile - f.tile_first[s]) * f.tile_bytes;
    const uint32_t tend = (uint32_t)jb_min((int)(tstart + f.tile_bytes), (int)len);
    const unsigned wpt = f.tile_bytes >> 5;
    uint32_t* B = f.vbits + (size_t)tile * wpt;
    for (unsigned j = 0; j < wpt; ++j) B[j] = 0u;
    JbWalker w;
    w.init(f.in + f.plane_off[s], len);
    // first offset of the tile that can start a block: 0, or the byte after a 0x00
    uint32_t pos = tstart;
    if (tstart != 0) {
        while (pos < tend && w.byte_at(pos - 1) != 0) ++pos;
    }
    uint32_t exit_pos = pos;                     // == tend if the tile holds no candidate
    if (pos < tend) {
        w.seek(pos);
        uint32_t last = pos;
        B[(pos - tstart) >> 5] |= 1u << ((pos - tstart) & 31u);
        int count = 0;
        // one flat loop over codes (not one loop per block): the 32 lanes of a warp stay converged
        // although their blocks end at different codes
        for (;;) {
            if (w.nb < 23) { w.buf = (w.buf << 32) | w.load(w.widx++); w.nb += 32; }
            const uint32_t head = (uint32_t)(w.buf >> (w.nb - 8)) & 0xFFu;
            const uint32_t size = head & 15u;
            const bool zero_size = size == 0u, eob = head == 0u;
            w.nb -= zero_size ? 8 : 8 + (int)size;
            count += zero_size ? (eob ? 0 : JB_MAX_RUN) : (int)(head >> 4) + 1;
            bool bad = (zero_size && !eob && head != 0xF0u) || size == 1u || count > f.n;
            uint32_t next = 0;
            if (eob) {
                uint32_t abs_bits = w.widx * 32u - (uint32_t)w.nb;
                const uint32_t pad = (8u - (abs_bits & 7u)) & 7u;
                w.nb -= (int)pad;
                abs_bits += pad;
                next = w.origin + ((abs_bits - w.shift_bits) >> 3);
                bad = bad || next * 8u > w.len_bits;
            }
            if (bad) {
                // the last recorded start was false (in a valid stream): resume at the next offset
                // that follows a 0x00 byte; the true chain joins the walk later
                next = last + 1u;
                while (next < tend && w.byte_at(next - 1) != 0) ++next;
                if (next >= tend) { exit_pos = JB_POS_INVALID; break; }
                w.seek(next);
            } else if (eob) {
                if (next >= tend) { exit_pos = next; break; }
            }
            if (bad || eob) {
                B[(next - tstart) >> 5] |= 1u << ((next - tstart) & 31u);
                last = next;
                count = 0;
            }
        }
    }
    f.tile_exit[tile] = exit_pos;
    jb_walk_note_exit(f, s, tile, tile - f.tile_first[s], f.tile_first[s + 1] - f.tile_first[s], exit_pos, len);
}

// A block that runs past the staged window of its walk: parse it from global memory.
// Returns the stream offset after the block, or JB_POS_INVALID.
__device__ __noinline__ uint32_t jb_walk_block_global(const uint8_t* stream, uint32_t len, uint32_t start, int n,
                                                      int maxblk) {
    JbWalker w;
    w.init(stream, len);
    w.seek(start);
    return w.block(n, (uint32_t)maxblk * 8u) ? (w.bp >> 3) : JB_POS_INVALID;
}

// Smallest q in [from, end) such that byte q-1 of the slice (big-endian-assembled words) is 0x00; `end` if none.
__device__ __forceinline__ uint32_t jb_next_candidate(const uint32_t* sl, uint32_t from, uint32_t end) {
    const uint32_t b = from - 1u;
    uint32_t wi = b >> 2;
    uint32_t v = sl[wi] | ~(0xFFFFFFFFu >> ((b & 3u) * 8u));        // bytes before b: made non-zero
    for (;;) {
        uint32_t t = (v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
        t = ~(t | v | 0x7F7F7F7Fu);                                 // 0x80 in every zero byte
        if (t) {
            const uint32_t q = wi * 4u + ((uint32_t)__clz((int)t) >> 3) + 1u;
            return q < end ? q : end;
        }
        ++wi;
        if (wi * 4u + 1u >= end) return end;
        v = sl[wi];
    }
}

// Small tiles.  A warp owns 32 consecutive tiles of ONE stream (f.warp_first: streams are padded to whole warps,
// so no warp spans two streams and the stream lookup happens once per warp): their bytes -- from the byte before
// the first tile to JB_WALK_HALO bytes behind the last -- are one contiguous run of the input, which the warp copies
// into shared memory with coalesced asynchronous copies (zero fill behind the stream's end), byte-swaps in place so
// that a funnel shift extracts any code head, and then every lane walks its own tile in it.  A block that leaves
// the tile is followed through the bytes of the next tiles (they are there anyway); only a block that leaves the
// warp's whole region is re-parsed from global memory.  Positions inside the walk are region-relative: region byte
// 0 is the 4-byte aligned address at or below the byte before the warp's first tile.  Behind the region: one
// bitmap of block starts per lane.
__global__ void __launch_bounds__(JB_WALK_THREADS) jb_frame_walk_smem_kernel(JbFrameArgs f, unsigned region_words,
                                                                             unsigned warp_words) {
    extern __shared__ uint32_t s_words[];
    const int lane = threadIdx.x & 31;
    const unsigned wpt = f.tile_bytes >> 5;
    const unsigned gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gwarp >= f.warp_first[f.n_planes]) return;                      // (whole warps)
    uint32_t* region = s_words + (size_t)(threadIdx.x >> 5) * warp_words;
    uint32_t* bm = region + region_words + (size_t)lane * wpt;
    const int s = (int)f.warp_stream[gwarp];                            // (written by the prep kernel)
    const unsigned t0 = (gwarp - f.warp_first[s]) * 32u;                // first tile of the warp, stream-relative
    const unsigned nt = f.tile_first[s + 1] - f.tile_first[s];
    const bool live = t0 + (unsigned)lane < nt;
    const unsigned tile = f.tile_first[s] + t0 + (unsigned)lane;
    const uint32_t len = (uint32_t)f.plane_len[s];
    const uint8_t* stream = f.in + f.plane_off[s];
    const uint32_t first = t0 ? t0 * f.tile_bytes - 1u : 0u;
    const uint32_t last = (uint32_t)jb_min((unsigned long long)(t0 + 32u) * f.tile_bytes + JB_WALK_HALO, (unsigned long long)len);
    const unsigned long long a0 = (unsigned long long)(uintptr_t)(stream + first);
    const uint32_t mis = (uint32_t)(a0 & 3ull);
    const uint32_t* gsrc = (const uint32_t*)(uintptr_t)(a0 - mis);
    uint32_t nw = (last - first + mis + 3u) >> 2;
    if (nw > region_words - 2u) nw = region_words - 2u;
    {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(region);
        for (uint32_t idx = lane; idx < region_words; idx += 32) {
            const bool in = idx < nw;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;"
                         :: "r"(dst + idx * 4u), "l"(in ? gsrc + idx : (const uint32_t*)f.tile_first), "r"(in ? 4 : 0) : "memory");
        }
        for (unsigned j = 0; j < wpt; ++j) bm[j] = 0u;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    for (uint32_t i = lane; i < region_words; i += 32) region[i] = jb_bswap32(region[i]);
    __syncwarp();
    if (live) {
        const uint32_t* mine = region;
        const uint32_t tstart = (t0 + (unsigned)lane) * f.tile_bytes;
        const uint32_t tend = (uint32_t)jb_min((unsigned long long)tstart + f.tile_bytes, (unsigned long long)len);
        const uint32_t to_stream = first - mis;             // region byte q is stream byte q + to_stream (mod 2^32)
        const uint32_t tstart_s = tstart - to_stream, tend_s = tend - to_stream, len_s = len - to_stream;
        const uint32_t staged_bits = nw * 32u;
        // first offset of the tile that can start a block: 0, or the byte after a 0x00
        uint32_t q = tstart ? jb_next_candidate(mine, tstart_s, tend_s) : tstart_s;
        uint32_t exit_pos = tend;                            // the tile holds no candidate
        if (q < tend_s) {
            uint32_t last_q = q;
            bm[(q - tstart_s) >> 5] |= 1u << ((q - tstart_s) & 31u);
            const uint32_t bm_addr = (uint32_t)__cvta_generic_to_shared(bm);
            uint32_t p = q * 8u;                             // bit position inside the region
            // One flat loop over codes, not one loop per block, and the end-of-block bookkeeping is
            // predicated: the 32 lanes of a warp stay converged although their blocks end at different
            // codes.  Only false starts, blocks that leave the staged region and the end of the tile branch.
            // (No coefficient count here: a parse that started on a false offset runs into an impossible
            // code within a few codes, or into the zero fill behind the region; a parse that started on a
            // true offset of a valid stream never exceeds the count.  The transform kernel checks it.)
            // a block that ends at or beyond this bit leaves the tile
            const uint32_t lim_bits = tend_s * 8u - 8u;
            const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(mine);
            for (;;) {
                // Two codes per pass from one 32-bit window: a code is at most 8 + 15 bits long, so the head of
                // the code behind it lies inside the window too.  The second code counts only if the first one
                // neither ends the block (the next block starts on a byte boundary) nor is impossible.
                const uint32_t wa = s_base + ((p >> 3) & ~3u);
                uint32_t w0, w1;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(wa));
                asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(wa));
                const uint32_t x1 = __funnelshift_l(w1, w0, p);
                // "stop": the code ends the block (EOB) or is impossible ((r, 0) with 0 < r < 15, or size 1)
                const uint32_t a1 = 8u + ((x1 >> 24) & 15u);
                const bool stop1 = (x1 & 0x0E000000u) == 0u && (x1 & 0xFF000000u) != 0xF0000000u;
                const uint32_t x2 = x1 << a1;
                const uint32_t xs = stop1 ? x1 : x2;                 // the code that decides this pass
                const uint32_t a2 = stop1 ? 0u : 8u + ((x2 >> 24) & 15u);
                p += a1 + a2;
                const bool stop = (xs & 0x0E000000u) == 0u && (xs & 0xFF000000u) != 0xF0000000u;
                q = (p + 7u) >> 3;
                if (p > staged_bits || (stop && (xs >= 0x01000000u || p > lim_bits))) {
                    bool false_start = stop && xs >= 0x01000000u;
                    if (p > staged_bits) {
                        // the parse left the staged region (or ran into the zero fill behind the stream):
                        // repeat this block from global memory
                        const uint32_t nx = jb_walk_block_global(stream, len, last_q + to_stream, f.n, f.maxblk);
                        false_start = nx == JB_POS_INVALID;
                        q = nx - to_stream;
                    } else if (q > len_s) {
                        false_start = true;
                    }
                    if (false_start) {
                        // the last recorded start was false (in a valid stream): resume at the next offset
                        // that follows a 0x00 byte; the true chain joins the walk later
                        q = jb_next_candidate(mine, last_q + 1u, tend_s);
                        if (q >= tend_s) { exit_pos = JB_POS_INVALID; break; }
                    } else if (q >= tend_s) {
                        exit_pos = q + to_stream;
                        break;
                    }
                    // q is the next start to record
                    const uint32_t rel = q - tstart_s;
                    asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(bm_addr + ((rel >> 5) << 2)), "r"(1u << (rel & 31u)) : "memory");
                    last_q = q;
                    p = q * 8u;
                } else if (stop) {
                    // a block ended inside the tile: q starts the next one.  (A branch, not a predicated update: measured
                    // 155 us against 162 us for the whole walk, and q of a pass without a block end may lie anywhere.)
                    const uint32_t rel = q - tstart_s;
                    asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(bm_addr + ((rel >> 5) << 2)), "r"(1u << (rel & 31u)) : "memory");
                    last_q = q;
                    p = q * 8u;
                }
            }
        }
        f.tile_exit[tile] = exit_pos;
        jb_walk_note_exit(f, s, tile, t0 + (unsigned)lane, nt, exit_pos, len);
    }
    __syncwarp();
    // the warp's bitmaps are consecutive in shared and in global memory
    {
        const unsigned tile0 = f.tile_first[s] + t0;
        const unsigned nlive = jb_min(32u, nt - t0);
        for (unsigned i = lane; i < nlive * wpt; i += 32u) f.vbits[(size_t)tile0 * wpt + i] = region[region_words + i];
    }
}

// ---- F2: chain of tiles per stream (pointer doubling over "tile of my exit") -----------------------
// CTA-wide.  J0/J1: uint16 [cap], R: uint8 [cap] in shared memory; nt <= cap.  Marks the stream for the
// serial walk if the chain does not end exactly at the end of the stream.
__device__ __forceinline__ void jb_reach_stream(const JbFrameArgs& f, int s, uint16_t* J0, uint16_t* J1, uint8_t* R) {
    uint16_t* J[2] = {J0, J1};
    const int tid = threadIdx.x;
    const unsigned t0 = f.tile_first[s];
    const unsigned nt = f.tile_first[s + 1] - t0;
    const uint32_t len = (uint32_t)f.plane_len[s];
    const unsigned T = f.tile_bytes;
    // (tile sizes are powers of two; the loops take four tiles per thread and pass so that their global loads overlap:
    // a long stream is 16 K tiles for one CTA, and a loop of dependent L2 round trips was most of this kernel)
    const unsigned tsh = (unsigned)__ffs((int)T) - 1u;
    const unsigned bd = blockDim.x;
    // every walk left its tile for the next one and the last one ended at the end of the stream: the walks have written
    // the entries already (jb_walk_note_exit)
    if (f.notplain[s] == 0u) return;
    for (unsigned base = tid; base < nt; base += 4u * bd) {
        uint32_t ex[4];
        #pragma unroll
        for (int k = 0; k < 4; ++k) ex[k] = base + k * bd < nt ? __ldg(f.tile_exit + t0 + base + k * bd) : JB_POS_INVALID;
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned t = base + k * bd;
            if (t >= nt) break;
            uint16_t nx = 0xFFFFu;
            if (ex[k] != JB_POS_INVALID && ex[k] < len) { const unsigned kk = ex[k] >> tsh; if (kk > t && kk < nt) nx = (uint16_t)kk; }
            J[0][t] = nx;
            R[t] = t == 0 ? 1 : 0;
            f.tile_entry[t0 + t] = t == 0 ? 0u : JB_POS_INVALID;
        }
    }
    // The usual stream: every walk leaves its tile for the next one (a block longer than a whole tile is what it takes
    // to skip one), i.e. every tile is on the chain -- seen in one parallel pass.  Pointer doubling (log2(tiles) rounds
    // by one CTA: 60 us for the 16 K tiles of a gigapixel plane) is for the streams where that does not hold.
    __syncthreads();
    int plain = 1;
    for (unsigned t = tid; t + 1u < nt; t += bd) plain &= J[0][t] == (uint16_t)(t + 1u);
    plain = __syncthreads_and(plain);
    if (plain) {
        for (unsigned t = tid; t < nt; t += bd) R[t] = 1;
        __syncthreads();
    }
    int cur = 0;
    for (unsigned span = 1; span < nt && !plain; span <<= 1) {
        for (unsigned base = tid; base < nt; base += 4u * bd) {
            uint16_t j[4], jj[4];
            #pragma unroll
            for (int k = 0; k < 4; ++k) j[k] = base + k * bd < nt ? J[cur][base + k * bd] : (uint16_t)0xFFFFu;
            #pragma unroll
            for (int k = 0; k < 4; ++k) jj[k] = j[k] != 0xFFFFu ? J[cur][j[k]] : (uint16_t)0xFFFFu;
            #pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned t = base + k * bd;
                if (t >= nt) break;
                if (j[k] != 0xFFFFu && R[t]) R[j[k]] = 1;        // marks only tiles that are on the chain
                J[cur ^ 1][t] = jj[k];
            }
        }
        __syncthreads();
        cur ^= 1;
    }
    // entries of the tiles on the chain, and the end of the chain
    int bad = 0, ends = 0;
    for (unsigned base = tid; base < nt; base += 4u * bd) {
        uint32_t ex[4];
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned t = base + k * bd;
            ex[k] = (t < nt && R[t]) ? __ldg(f.tile_exit + t0 + t) : 0xFFFFFFFEu;          // 0xFFFFFFFE: not on the chain
        }
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned t = base + k * bd;
            if (ex[k] == 0xFFFFFFFEu) continue;
            if (ex[k] == len) ++ends;
            else if (ex[k] == JB_POS_INVALID || ex[k] > len || (ex[k] >> tsh) <= t || (ex[k] >> tsh) >= nt) bad = 1;
            else f.tile_entry[t0 + (ex[k] >> tsh)] = ex[k];
        }
    }
    bad = __syncthreads_or(bad);
    const int total_ends = __syncthreads_count(ends);
    if (tid == 0 && (bad || total_ends != 1)) f.fallback[s] = 1u;
    __syncthreads();
}

// CAP = tiles of one stream this instantiation handles in shared memory; longer streams are left to
// the larger instantiation (or, beyond that, to the serial fallback).
template <unsigned CAP, unsigned MIN_TILES>
__global__ void __launch_bounds__(1024) jb_frame_reach_kernel(JbFrameArgs f) {
    extern __shared__ __align__(16) unsigned char reach_smem[];
    const int tid = threadIdx.x;
    if (f.tile_first[f.n_planes] == 0) return;
    // small instantiation: one CTA per stream; large one: a few CTAs share the list of long streams
    const unsigned n_work = MIN_TILES ? f.big_list[0] : (unsigned)f.n_planes;
    for (unsigned wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        const int s = MIN_TILES ? (int)f.big_list[1 + wi] : (int)wi;
        const unsigned nt = f.tile_first[s + 1] - f.tile_first[s];
        if (nt < MIN_TILES) continue;                            // the smaller instantiation took it
        if (nt != 0 && f.notplain[s] == 0u) continue;            // every tile on the chain: nothing to work out, however long
        if (nt > CAP) { if (CAP >= 40000u && tid == 0) f.fallback[s] = 1u; continue; }
        if (nt == 0) { if (tid == 0) f.fallback[s] = 1u; continue; }
        jb_reach_stream(f, s, (uint16_t*)reach_smem, (uint16_t*)reach_smem + CAP, (uint8_t*)(reach_smem + 4 * (size_t)CAP));
    }
}

// ---- F2a: link a tile on the chain to its entry ----------------------------------------------------
__device__ __forceinline__ void jb_link_tile(const JbFrameArgs& f, int s, unsigned tile) {
    const uint32_t len = (uint32_t)f.plane_len[s];
    const unsigned t = tile - f.tile_first[s];
    const uint32_t tstart = t * f.tile_bytes;
    const uint32_t tend = (uint32_t)jb_min((int)(tstart + f.tile_bytes), (int)len);
    const unsigned wpt = f.tile_bytes >> 5;
    const uint32_t* B = f.vbits + (size_t)tile * wpt;
    const uint32_t my_exit = f.tile_exit[tile];

    const uint32_t E = f.tile_entry[tile];                  // JB_POS_INVALID: no true block starts here
    unsigned from = f.tile_bytes, npriv = 0, hops = 0;      // from: byte of the tile where the chain joins the walk
    bool ok = true;
    if (E == JB_POS_INVALID) {
        // not on the chain of tiles: a block that started earlier covers this tile
    } else if (E < tstart || E >= tend || my_exit == JB_POS_INVALID) {
        ok = false;
    } else {
        const uint32_t rel = E - tstart;
        if ((B[rel >> 5] >> (rel & 31u)) & 1u) {
            from = rel;
        } else {
            // follow the chain from E until it meets the walk (or leaves the tile at the walk's exit)
            JbWalker w;
            w.init(f.in + f.plane_off[s], len);
            w.seek(E);
            const uint32_t maxblk_bits = (uint32_t)f.maxblk * 8u;
            for (;;) {
                if (!w.block(f.n, maxblk_bits)) { ok = false; break; }
                ++npriv;
                const uint32_t pos = w.bp >> 3;
                if (pos >= tend) { ok = (pos == my_exit); break; }
                const uint32_t r = pos - tstart;
                if ((B[r >> 5] >> (r & 31u)) & 1u) { from = r; break; }
            }
        }
        hops = npriv;
        if (from < f.tile_bytes) {
            hops += (unsigned)__popc(B[from >> 5] >> (from & 31u));
            for (unsigned j = (from >> 5) + 1u; j < wpt; ++j) hops += (unsigned)__popc(B[j]);
        }
    }
    if (!ok) { f.fallback[s] = 1u; hops = 0; from = f.tile_bytes; npriv = 0; }
    f.tile_from[tile] = from;
    f.tile_npriv[tile] = npriv;
    f.tile_hops[tile] = hops;
}

__global__ void __launch_bounds__(JB_FRAME_THREADS) jb_frame_link_kernel(JbFrameArgs f) {
    const unsigned tile = blockIdx.x * JB_FRAME_THREADS + threadIdx.x;
    if (tile >= f.tile_first[f.n_planes]) return;
    jb_link_tile(f, jb_stream_of_tile(f.tile_first, f.n_planes, tile), tile);
}

// ---- F2b: per stream, ordinals of the tiles' first blocks (CTA-wide) -----------------------------------
template <int PER>
__device__ __forceinline__ void jb_scan_stream(const JbFrameArgs& f, int s, unsigned* s_warp) {
    const int tid = threadIdx.x;
    const unsigned t0 = f.tile_first[s];
    const unsigned nt = f.tile_first[s + 1] - t0;
    unsigned carry = 0;
    // PER consecutive tiles per thread and pass: one block scan (and one round of dependent L2 round trips) per
    // PER x blockDim tiles (PER = 4; 20 -- a gigapixel plane's 16 K tiles in one pass of 1024 threads -- measured slower,
    // 14.4 against 12.0 us: the loads of neighbouring threads no longer share sectors)
    for (unsigned b = 0; b < nt; b += (unsigned)PER * blockDim.x) {
        const unsigned t = b + (unsigned)PER * (unsigned)tid;
        unsigned h[PER];
        unsigned sum = 0;
        #pragma unroll
        for (int k = 0; k < PER; ++k) { h[k] = t + k < nt ? f.tile_hops[t0 + t + k] : 0u; sum += h[k]; }
        unsigned total;
        unsigned ex = carry + jb_block_excl_scan(sum, s_warp, &total);
        #pragma unroll
        for (int k = 0; k < PER; ++k) {
            if (t + k < nt) f.tile_base[t0 + t + k] = ex;
            ex += h[k];
        }
        carry += total;
    }
    if (tid == 0) {
        // (that the chain of tiles ends exactly at the stream end was checked by jb_reach_stream)
        const bool good = nt > 0 && f.fallback[s] == 0u && carry == (unsigned)f.nblocks;
        if (!good) {
            f.fallback[s] = 1u;
            atomicAdd(f.status + 2, 1ull);        // status[2]: streams that took the serial walk
        }
    }
}

__global__ void __launch_bounds__(1024) jb_frame_scan_kernel(JbFrameArgs f) {
    __shared__ unsigned s_warp[33];
    if (f.tile_first[f.n_planes] == 0) return;                // prep failed, error already set
    jb_scan_stream<4>(f, blockIdx.x, s_warp);
}

// ---- F3: emit the block offsets of a tile -------------------------------------------------------------
__device__ __forceinline__ void jb_emit_tile(const JbFrameArgs& f, int s, unsigned tile) {
    const uint32_t len = (uint32_t)f.plane_len[s];
    const uint32_t tstart = (tile - f.tile_first[s]) * f.tile_bytes;
    const unsigned wpt = f.tile_bytes >> 5;
    const uint32_t* B = f.vbits + (size_t)tile * wpt;
    const unsigned from = f.tile_from[tile], npriv = f.tile_npriv[tile];
    unsigned idx = f.tile_base[tile];
    unsigned* out = f.block_start + (size_t)s * f.nblocks;
    if (npriv) {
        JbWalker w;
        w.init(f.in + f.plane_off[s], len);
        w.seek(f.tile_entry[tile]);
        const uint32_t maxblk_bits = (uint32_t)f.maxblk * 8u;
        for (unsigned k = 0; k < npriv && idx < (unsigned)f.nblocks; ++k) {
            out[idx++] = w.bp >> 3;
            w.block(f.n, maxblk_bits);
        }
    }
    if (from < f.tile_bytes) {
        if (wpt == 8u || wpt == 16u) {
            // the usual tiles (256 or 512 bytes): the whole bitmap in two or four 128-bit loads, then registers only --
            // word by word from memory, every word of the walk below was another dependent L2 round trip
            // (59 % of the stitch kernel's stall samples)
            const uint4* B4 = (const uint4*)B;                   // (tile bitmaps are 32 / 64 bytes apart from a 256-byte base)
            uint32_t bw[16];
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if ((unsigned)q * 4u < wpt) v = B4[q];
                bw[4 * q] = v.x; bw[4 * q + 1] = v.y; bw[4 * q + 2] = v.z; bw[4 * q + 3] = v.w;
            }
            const unsigned j0 = from >> 5;
            #pragma unroll
            for (unsigned j = 0; j < 16u; ++j) {
                if (j < j0 || j >= wpt) continue;
                uint32_t word = bw[j];
                if (j == j0) word &= 0xFFFFFFFFu << (from & 31u);
                while (word && idx < (unsigned)f.nblocks) {
                    out[idx++] = tstart + j * 32u + (unsigned)__ffs((int)word) - 1u;
                    word &= word - 1u;
                }
            }
        } else {
            uint32_t word = B[from >> 5] & (0xFFFFFFFFu << (from & 31u));
            for (unsigned j = from >> 5;;) {
                while (word && idx < (unsigned)f.nblocks) {
                    out[idx++] = tstart + j * 32u + (unsigned)__ffs((int)word) - 1u;
                    word &= word - 1u;
                }
                if (++j >= wpt) break;
                word = B[j];
            }
        }
    }
}

__global__ void __launch_bounds__(JB_FRAME_THREADS) jb_frame_emit_kernel(JbFrameArgs f) {
    const unsigned tile = blockIdx.x * JB_FRAME_THREADS + threadIdx.x;
    if (tile >= f.tile_first[f.n_planes]) return;
    const int s = jb_stream_of_tile(f.tile_first, f.n_planes, tile);
    if (f.fallback[s]) return;
    jb_emit_tile(f, s, tile);
}

// ---- F4: serial fallback, one warp per marked stream ------------------------------------------------------
// The warp stages the stream through shared memory 4 KB at a time (coalesced); lane 0 walks it.
__device__ __forceinline__ void jb_serial_stream(const JbFrameArgs& f, int s, uint32_t* sw, int lane) {
    const uint32_t len = (uint32_t)f.plane_len[s];
    const uint8_t* stream = f.in + f.plane_off[s];
    unsigned* out = f.block_start + (size_t)s * f.nblocks;
    const uint32_t maxblk_bits = (uint32_t)f.maxblk * 8u;
    uint32_t pos = 0;
    unsigned k = 0;
    int ok = (len > 0 && f.tile_first[f.n_planes] != 0) ? 1 : 0;
    while (ok && pos < len) {
        const unsigned long long a0 = (unsigned long long)(uintptr_t)(stream + pos);
        const uint32_t mis = (uint32_t)(a0 & 3ull);
        const uint32_t* g = (const uint32_t*)(uintptr_t)(a0 - mis);
        const uint32_t gnw = (len - pos + mis + 3u) >> 2;
        const uint32_t nw = gnw < JB_SERIAL_WORDS ? gnw : JB_SERIAL_WORDS;
        for (uint32_t i = lane; i < nw; i += 32) sw[i] = __ldg(g + i);
        __syncwarp();
        if (lane == 0) {
            JbWalker w;
            w.init_window(sw, nw, g, gnw, pos, mis, len);
            w.seek(pos);
            const uint32_t stop = pos + (JB_SERIAL_WORDS - 64) * 4u;       // blocks that start before this
            while (ok && (w.bp >> 3) < len && (w.bp >> 3) < stop) {
                if (k >= (unsigned)f.nblocks) { ok = 0; break; }
                out[k++] = w.bp >> 3;
                ok = w.block(f.n, maxblk_bits) ? 1 : 0;
            }
            pos = w.bp >> 3;
        }
        pos = __shfl_sync(0xffffffffu, pos, 0);
        ok = __shfl_sync(0xffffffffu, ok, 0);
        k = __shfl_sync(0xffffffffu, k, 0);
        __syncwarp();
    }
    if (lane == 0 && (!ok || k != (unsigned)f.nblocks)) jb_set_error(f.status, JB_ERR_BAD_STREAM);
}

__global__ void __launch_bounds__(32) jb_frame_serial_kernel(JbFrameArgs f) {
    __shared__ uint32_t sw[JB_SERIAL_WORDS + 8];
    if (f.fallback[blockIdx.x] == 0u) return;
    jb_serial_stream(f, blockIdx.x, sw, threadIdx.x);
}

// ---- F2..F4 in one launch, for batches of short streams: one CTA per stream ---------------------------------
// (the per-tile arrays written by one phase and read by the next stay in L2; __syncthreads orders them)
#define JB_STITCH_THREADS 256
__global__ void __launch_bounds__(JB_STITCH_THREADS) jb_frame_stitch_kernel(JbFrameArgs f) {
    extern __shared__ __align__(16) unsigned char stitch_smem[];       // max(5 * cap, serial window)
    __shared__ unsigned s_warp[33];
    const int s = blockIdx.x, tid = threadIdx.x;
    if (f.tile_first[f.n_planes] == 0) return;                // prep failed, error already set
    const unsigned t0 = f.tile_first[s];
    const unsigned nt = f.tile_first[s + 1] - t0;
    const unsigned cap = f.stitch_cap;
    if (nt == 0 || (nt > cap && f.notplain[s] != 0u)) {
        if (tid == 0) f.fallback[s] = 1u;                     // empty (invalid), or off the usual chain and far longer than its peers
        __syncthreads();
    } else {
        jb_reach_stream(f, s, (uint16_t*)stitch_smem, (uint16_t*)stitch_smem + cap, (uint8_t*)(stitch_smem + 4 * (size_t)cap));
        for (unsigned t = tid; t < nt; t += blockDim.x) jb_link_tile(f, s, t0 + t);
        __syncthreads();
    }
    jb_scan_stream<4>(f, s, s_warp);
    __syncthreads();
    if (f.fallback[s] == 0u) {
        for (unsigned t = tid; t < nt; t += blockDim.x) jb_emit_tile(f, s, t0 + t);
    } else if (tid < 32) {
        jb_serial_stream(f, s, (uint32_t*)stitch_smem, tid);
    }
}

cudaError_t jb_launch_framing(const JbFrameArgs& f_in, cudaStream_t s) {
    cudaError_t e;
    JbFrameArgs f = f_in;
    const unsigned grid = (f.max_tiles + JB_FRAME_THREADS - 1) / JB_FRAME_THREADS;
    JB_LAUNCH((jb_frame_prep_kernel), 1, 1024, 0, s, f);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    // region of one warp of the walk: 32 tiles, the byte before them, JB_WALK_HALO bytes behind them, misalignment and
    // two words of zero slack; then 32 bitmaps.  The block size that puts most warps on an SM.
    const unsigned region_words = 8u * (unsigned)f.tile_bytes + JB_WALK_HALO / 4u + 4u;
    const unsigned warp_words = region_words + (unsigned)f.tile_bytes;          // 32 bitmaps of tile_bytes / 32 words
    unsigned wthreads = 0, best_warps = 0;
    for (unsigned wt = JB_WALK_THREADS; wt >= 32u; wt -= 32u) {
        const size_t per_block = (size_t)(wt / 32u) * warp_words * 4u + 1024u;
        unsigned blocks = (unsigned)((size_t)227 * 1024 / per_block);
        if (blocks > 32u) blocks = 32u;
        const unsigned warps = blocks * wt / 32u;
        if (warps > best_warps && warps >= 8u) { best_warps = warps; wthreads = wt; }
    }
    if (wthreads) {
        const size_t smem = (size_t)(wthreads / 32u) * warp_words * 4;
        // every stream gets whole warps: at most max_tiles / 32 + n_planes of them
        const size_t walk_warps = (size_t)f.max_tiles / 32u + (size_t)f.n_planes + 1u;
        const size_t walk_blocks = (walk_warps * 32u + wthreads - 1) / wthreads;
        e = cudaFuncSetAttribute(jb_frame_walk_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        JB_LAUNCH((jb_frame_walk_smem_kernel), (unsigned)walk_blocks, wthreads, smem, s, f, region_words, warp_words);
    } else {
        JB_LAUNCH((jb_frame_walk_kernel), (f.max_tiles + JB_WALK_THREADS - 1) / JB_WALK_THREADS, JB_WALK_THREADS, 0, s, f);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (!jb_framing_is_chain(f.max_tiles, f.n_planes)) {
        // many short streams (a batch of images): the rest of the framing in one launch, one CTA per stream;
        // a stream of more than `cap` tiles (far longer than its peers) takes the serial walk
        // The kernel is a chain of short dependent phases (latency bound): streams of a few hundred tiles get
        // narrower CTAs and half the table space, so that twice as many of them are resident per SM.
        // (not when every stream already has its own CTA of full width in a single wave: 148 SMs x 8)
        const bool small = (size_t)f.max_tiles <= (size_t)f.n_planes * 256u && f.n_planes > 148 * 8;
        const unsigned cap_max = small ? JB_STITCH_CAP / 2u : JB_STITCH_CAP;
        unsigned cap = f.max_tiles < cap_max ? f.max_tiles : cap_max;
        cap = (cap + 7u) & ~7u;
        size_t smem = 5 * (size_t)cap;
        if (smem < (JB_SERIAL_WORDS + 8) * 4) smem = (JB_SERIAL_WORDS + 8) * 4;
        f.stitch_cap = cap;
        JB_LAUNCH((jb_frame_stitch_kernel), f.n_planes, small ? JB_STITCH_THREADS / 2 : JB_STITCH_THREADS, smem, s, f);
        return cudaGetLastError();
    }
    {   // chain of tiles: streams of up to 4096 tiles (1 MB) in 20 KB of shared memory, longer ones in 200 KB
        const size_t sm_small = 5 * 4096, sm_big = 5 * 40000;
        JB_LAUNCH((jb_frame_reach_kernel<JB_REACH_SMALL_CAP, 0>), f.n_planes, JB_FRAME_THREADS, sm_small, s, f);
        e = cudaFuncSetAttribute(jb_frame_reach_kernel<40000, JB_REACH_SMALL_CAP + 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_big);
        if (e != cudaSuccess) return e;
        JB_LAUNCH((jb_frame_reach_kernel<40000, JB_REACH_SMALL_CAP + 1>), f.n_planes < 32 ? f.n_planes : 32, 1024, sm_big, s, f);
    }
    JB_LAUNCH((jb_frame_link_kernel), grid, JB_FRAME_THREADS, 0, s, f);
    // a few long streams: wide CTAs; many short streams: narrow ones
    JB_LAUNCH((jb_frame_scan_kernel), f.n_planes, (f.max_tiles / (unsigned)f.n_planes > 2048u) ? 1024 : JB_FRAME_THREADS, 0, s, f);
    JB_LAUNCH((jb_frame_emit_kernel), grid, JB_FRAME_THREADS, 0, s, f);
    JB_LAUNCH((jb_frame_serial_kernel), f.n_planes, 32, 0, s, f);
    return cudaGetLastError();
}
