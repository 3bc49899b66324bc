// jb_framing.cu -- parallel discovery of block boundaries in RleBytestream output.
//
// The container stores one length per band and nothing else (file_format.py:86-93); a
// block ends wherever its EOB lands plus zero padding to the byte (rle_byte_stream.py:55-56,
// 80-82), so the reference finds block k only by decoding blocks 0..k-1 (BitDecoder,
// rle_byte_stream.py:6-42).  Here:
//
//   * every block ends with a 0x00 byte (the EOB's eight zero bits reach the end of a byte
//     and padding is zero), so a block can only start at offset 0 or after a 0x00 byte:
//     those offsets are the CANDIDATES;
//   * each stream is cut into tiles of JB_TILE_BYTES; per tile all candidates are parsed
//     in parallel (block extent only), linked to the candidate where their block ends,
//     and ranked by pointer jumping in shared memory -> for every candidate: the number
//     of blocks to the tile end and the offset where the chain leaves the tile;
//   * the first true block start of tile t+1 is the exit of tile t's true entry.  False
//     candidates re-synchronise with the true chain almost at once, so usually every
//     possible entry of a tile has the same exit and tile t+1's entry is known without
//     walking the tiles in order; only tiles where that fails are resolved serially;
//   * a scan over per-tile block counts gives every true block its ordinal, and a last
//     pass writes block_start[ordinal].
#include "jb_common.cuh"
#include "jb_inverse.cuh"

#define JB_TERM 0xFFFFu
#define JB_EX_INVALID 0xFFFFu

// ---- F0: tiles per stream -------------------------------------------------------------------
__global__ void __launch_bounds__(1024) jb_frame_prep_kernel(JbFrameArgs f) {
    __shared__ unsigned s_warp[33];
    unsigned carry = 0;
    bool bad = false;
    for (int base = 0; base < f.n_planes; base += 1024) {
        int s = base + threadIdx.x;
        unsigned nt = 0;
        if (s < f.n_planes) {
            unsigned long long len = f.plane_len[s];
            if (len > 0xFFFF0000ull) { bad = true; len = 0; }
            nt = (unsigned)((len + JB_TILE_BYTES - 1) / JB_TILE_BYTES);
        }
        unsigned total;
        unsigned ex = jb_block_excl_scan(nt, s_warp, &total);
        if (s < f.n_planes) f.tile_first[s] = carry + ex;
        carry += total;
    }
    if (__syncthreads_or(bad ? 1 : 0)) carry = 0xFFFFFFFFu;
    if (threadIdx.x == 0) {
        if (carry > f.max_tiles) {
            jb_set_error(f.status, JB_ERR_BAD_STREAM);
            carry = 0;
        }
        f.tile_first[f.n_planes] = carry;
    }
}

__device__ __forceinline__ int jb_stream_of_tile(const unsigned* tile_first, int n_planes, unsigned tile) {
    int lo = 0, hi = n_planes;                 // largest s < n_planes with tile_first[s] <= tile
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (tile_first[mid] <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- F1: per tile candidate parse + ranking -------------------------------------------------
struct JbTileSmem {
    size_t bytes, bitmap, wpre, pos, j0, j1, h0, h1, e0, e1, total;
};
__host__ __device__ inline JbTileSmem jb_tile_smem(int maxblk) {
    JbTileSmem L;
    size_t o = 0;
    L.bytes = o;  o += jb_align_up((size_t)JB_TILE_BYTES + maxblk + 16, 16);
    L.bitmap = o; o += (JB_TILE_BYTES / 32) * 4;
    L.wpre = o;   o += (JB_TILE_BYTES / 32) * 4;
    L.pos = o;    o += JB_TILE_BYTES * 2;
    L.j0 = o;     o += JB_TILE_BYTES * 2;
    L.j1 = o;     o += JB_TILE_BYTES * 2;
    L.h0 = o;     o += JB_TILE_BYTES * 2;
    L.h1 = o;     o += JB_TILE_BYTES * 2;
    L.e0 = o;     o += JB_TILE_BYTES * 2;
    L.e1 = o;     o += JB_TILE_BYTES * 2;
    L.total = o;
    return L;
}

__device__ __forceinline__ unsigned jb_cand_index(const unsigned* bitmap, const unsigned* wpre, unsigned p) {
    return wpre[p >> 5] + __popc(bitmap[p >> 5] & ((1u << (p & 31u)) - 1u));
}

__global__ void __launch_bounds__(JB_FRAME_THREADS) jb_frame_tiles_kernel(JbFrameArgs f) {
    extern __shared__ __align__(16) unsigned char smem[];
    const unsigned total_tiles = f.tile_first[f.n_planes];
    const unsigned tile = blockIdx.x;
    if (tile >= total_tiles) return;
    const JbTileSmem L = jb_tile_smem(f.maxblk);
    uint8_t* sBytes = smem + L.bytes;                 // sBytes[0] = byte before the tile
    unsigned* sBitmap = (unsigned*)(smem + L.bitmap);
    unsigned* sWpre = (unsigned*)(smem + L.wpre);
    uint16_t* sPos = (uint16_t*)(smem + L.pos);
    uint16_t* sJ[2] = {(uint16_t*)(smem + L.j0), (uint16_t*)(smem + L.j1)};
    uint16_t* sH[2] = {(uint16_t*)(smem + L.h0), (uint16_t*)(smem + L.h1)};
    uint16_t* sE[2] = {(uint16_t*)(smem + L.e0), (uint16_t*)(smem + L.e1)};
    __shared__ unsigned s_warp[33];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = jb_stream_of_tile(f.tile_first, f.n_planes, tile);
    const unsigned long long len = f.plane_len[s];
    const unsigned tstart = (tile - f.tile_first[s]) * JB_TILE_BYTES;
    const uint8_t* src = f.in + f.plane_off[s] + tstart;
    const unsigned remain = (unsigned)(len - tstart);
    const unsigned tb_eff = remain < JB_TILE_BYTES ? remain : JB_TILE_BYTES;
    const unsigned avail = remain < (unsigned)(JB_TILE_BYTES + f.maxblk) ? remain : (unsigned)(JB_TILE_BYTES + f.maxblk);

    for (unsigned i = tid; i < avail + 1; i += JB_FRAME_THREADS)
        sBytes[i] = (i == 0) ? (tstart == 0 ? (uint8_t)0 : src[-1]) : src[i - 1];
    __syncthreads();

    // candidate bitmap: offset p may start a block iff the byte before it is 0x00
    for (int w = warp; w < JB_TILE_BYTES / 32; w += JB_FRAME_THREADS / 32) {
        unsigned p = w * 32 + lane;
        unsigned m = __ballot_sync(0xffffffffu, p < tb_eff && sBytes[p] == 0);
        if (lane == 0) sBitmap[w] = m;
    }
    __syncthreads();
    unsigned ncand;
    {
        unsigned c = tid < JB_TILE_BYTES / 32 ? __popc(sBitmap[tid]) : 0u;
        unsigned ex = jb_block_excl_scan(c, s_warp, &ncand);
        if (tid < JB_TILE_BYTES / 32) sWpre[tid] = ex;
    }
    __syncthreads();
    for (unsigned p = tid; p < tb_eff; p += JB_FRAME_THREADS)
        if (sBitmap[p >> 5] >> (p & 31u) & 1u) sPos[jb_cand_index(sBitmap, sWpre, p)] = (uint16_t)p;
    __syncthreads();

    // parse the extent of the block that would start at each candidate
    for (unsigned c = tid; c < ncand; c += JB_FRAME_THREADS) {
        unsigned p = sPos[c], e = 0;
        uint16_t J = JB_TERM, Hh = 0, E = JB_EX_INVALID;
        if (jb_parse_block_extent(sBytes + 1, p, avail, f.n, (uint32_t)f.maxblk, &e) == JB_PARSE_OK) {
            if (e >= tb_eff) { Hh = 1; E = (uint16_t)e; }
            else if (sBitmap[e >> 5] >> (e & 31u) & 1u) { Hh = 1; J = (uint16_t)jb_cand_index(sBitmap, sWpre, e); }
        }
        sJ[0][c] = J; sH[0][c] = Hh; sE[0][c] = E;
        f.cand_pos[(size_t)tile * JB_TILE_BYTES + c] = (uint16_t)p;
        f.cand_next[(size_t)tile * JB_TILE_BYTES + c] = J;
    }
    __syncthreads();

    // pointer jumping: hops to, and offset of, the point where each chain leaves the tile
    int cur = 0;
    for (unsigned span = 1; span < ncand; span <<= 1) {
        const int nxt = cur ^ 1;
        for (unsigned c = tid; c < ncand; c += JB_FRAME_THREADS) {
            uint16_t j = sJ[cur][c];
            uint16_t hh = sH[cur][c], ee = sE[cur][c], jj = j;
            if (j != JB_TERM) {
                hh = (uint16_t)(hh + sH[cur][j]);
                jj = sJ[cur][j];
                if (jj == JB_TERM) ee = sE[cur][j];
            }
            sJ[nxt][c] = jj; sH[nxt][c] = hh; sE[nxt][c] = ee;
        }
        __syncthreads();
        cur = nxt;
    }

    // window table + "all entry candidates agree on the exit"
    const unsigned ref = (ncand > 0 && sPos[0] < f.win_n) ? sE[cur][0] : JB_EX_INVALID;
    int agree = 1;
    for (unsigned p = tid; p < f.win_n; p += JB_FRAME_THREADS) {
        unsigned val = JB_U32_NONE;
        if (p < tb_eff && (sBitmap[p >> 5] >> (p & 31u) & 1u)) {
            unsigned c = jb_cand_index(sBitmap, sWpre, p);
            unsigned ee = sE[cur][c], hh = sH[cur][c];
            val = (hh << 16) | ee;
            if (ee != ref) agree = 0;
        }
        f.win[(size_t)tile * f.win_n + p] = val;
    }
    agree = __syncthreads_and(agree);
    if (tid == 0) {
        f.tile_uniq[tile] = (agree && ref != JB_EX_INVALID) ? ref : JB_U32_NONE;
        f.tile_ncand[tile] = ncand;
    }
}

// ---- F2: per stream entry resolution + block ordinals ---------------------------------------
#define JB_UNK_CAP 256
__global__ void __launch_bounds__(JB_FRAME_THREADS) jb_frame_resolve_kernel(JbFrameArgs f) {
    __shared__ unsigned s_warp[33];
    __shared__ unsigned s_unk[JB_UNK_CAP];
    __shared__ unsigned s_nunk;
    const int s = blockIdx.x, tid = threadIdx.x;
    if (f.tile_first[f.n_planes] == 0) return;                // prep failed, error already set
    const unsigned t0 = f.tile_first[s];
    const unsigned nt = f.tile_first[s + 1] - t0;
    const unsigned long long len = f.plane_len[s];
    if (nt == 0) { if (tid == 0) jb_set_error(f.status, JB_ERR_BAD_STREAM); return; }
    if (tid == 0) s_nunk = 0;
    __syncthreads();

    // entries that follow from "every candidate of the previous tile exits at the same offset"
    for (unsigned t = tid; t < nt; t += JB_FRAME_THREADS) {
        unsigned e = 0;
        if (t > 0) {
            unsigned u = f.tile_uniq[t0 + t - 1];
            if (u != JB_U32_NONE && u >= JB_TILE_BYTES) e = u - JB_TILE_BYTES;
            else {
                e = JB_U32_NONE;
                unsigned k = atomicAdd(&s_nunk, 1u);
                if (k < JB_UNK_CAP) s_unk[k] = t;
            }
        }
        f.tile_entry[t0 + t] = e;
    }
    __syncthreads();
    // the rest, in tile order, through the window tables (rare)
    if (tid == 0 && s_nunk > 0) {
        const bool listed = s_nunk <= JB_UNK_CAP;
        if (listed) {                                           // ascending order
            for (unsigned i = 1; i < s_nunk; ++i) {
                unsigned v = s_unk[i]; int j = (int)i - 1;
                while (j >= 0 && s_unk[j] > v) { s_unk[j + 1] = s_unk[j]; --j; }
                s_unk[j + 1] = v;
            }
        }
        const unsigned count = listed ? s_nunk : nt - 1;
        for (unsigned i = 0; i < count; ++i) {
            unsigned t = listed ? s_unk[i] : i + 1;
            if (f.tile_entry[t0 + t] != JB_U32_NONE) continue;
            unsigned ep = f.tile_entry[t0 + t - 1];
            if (ep == JB_U32_NONE || ep >= f.win_n) break;
            unsigned w = f.win[(size_t)(t0 + t - 1) * f.win_n + ep];
            unsigned ex = w & 0xFFFFu;
            if (w == JB_U32_NONE || ex == JB_EX_INVALID || ex < JB_TILE_BYTES) break;
            f.tile_entry[t0 + t] = ex - JB_TILE_BYTES;
        }
    }
    __syncthreads();

    // hops per tile from the true entry, consistency checks, exclusive scan -> ordinals
    unsigned carry = 0;
    int bad = 0;
    for (unsigned b = 0; b < nt; b += JB_FRAME_THREADS) {
        unsigned t = b + tid, hops = 0;
        if (t < nt) {
            unsigned e = f.tile_entry[t0 + t];
            // the last tile may hold nothing but the tail of a block that started earlier
            const bool tail_only = (t + 1 == nt) && e != JB_U32_NONE && (unsigned long long)t * JB_TILE_BYTES + e == len;
            unsigned w = (!tail_only && e != JB_U32_NONE && e < f.win_n) ? f.win[(size_t)(t0 + t) * f.win_n + e] : JB_U32_NONE;
            unsigned ex = w & 0xFFFFu;
            if (tail_only) hops = 0;
            else if (w == JB_U32_NONE || ex == JB_EX_INVALID) bad = 1;
            else {
                hops = w >> 16;
                if (t + 1 < nt) {
                    if (ex < JB_TILE_BYTES || ex - JB_TILE_BYTES != f.tile_entry[t0 + t + 1]) bad = 1;
                } else if ((unsigned long long)t * JB_TILE_BYTES + ex != len) bad = 1;
            }
        }
        unsigned total;
        unsigned ex_scan = jb_block_excl_scan(hops, s_warp, &total);
        if (t < nt) { f.tile_base[t0 + t] = carry + ex_scan; f.tile_hops[t0 + t] = bad ? 0u : hops; }
        carry += total;
    }
    bad = __syncthreads_or(bad);
    if (tid == 0 && (bad || carry != (unsigned)f.nblocks)) jb_set_error(f.status, JB_ERR_BAD_STREAM);
}

// ---- F3: per tile, write the offset of every true block --------------------------------------
#define JB_EMIT_LEVELS 12
__global__ void __launch_bounds__(JB_FRAME_THREADS) jb_frame_emit_kernel(JbFrameArgs f) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint16_t* sPos = (uint16_t*)smem;
    uint16_t* sLvl = sPos + JB_TILE_BYTES;                    // [levels][JB_TILE_BYTES]
    __shared__ unsigned s_eidx;
    const unsigned total_tiles = f.tile_first[f.n_planes];
    const unsigned tile = blockIdx.x;
    if (tile >= total_tiles) return;
    const int tid = threadIdx.x;
    const unsigned hops = f.tile_hops[tile];
    const unsigned entry = f.tile_entry[tile];
    if (hops == 0 || entry == JB_U32_NONE) return;
    const int s = jb_stream_of_tile(f.tile_first, f.n_planes, tile);
    const unsigned tstart = (tile - f.tile_first[s]) * JB_TILE_BYTES;
    const unsigned ncand = f.tile_ncand[tile];
    const unsigned base = f.tile_base[tile];
    int levels = 0;
    while (levels < JB_EMIT_LEVELS && (1u << levels) < hops) ++levels;

    for (unsigned c = tid; c < ncand; c += JB_FRAME_THREADS) {
        sPos[c] = f.cand_pos[(size_t)tile * JB_TILE_BYTES + c];
        sLvl[c] = f.cand_next[(size_t)tile * JB_TILE_BYTES + c];
    }
    if (tid == 0) s_eidx = JB_U32_NONE;
    __syncthreads();
    for (int k = 1; k < levels; ++k) {
        const uint16_t* prev = sLvl + (size_t)(k - 1) * JB_TILE_BYTES;
        uint16_t* cur = sLvl + (size_t)k * JB_TILE_BYTES;
        for (unsigned c = tid; c < ncand; c += JB_FRAME_THREADS) {
            uint16_t j = prev[c];
            cur[c] = (j == JB_TERM) ? (uint16_t)JB_TERM : prev[j];
        }
        __syncthreads();
    }
    if (tid == 0) {                                           // index of the entry among the candidates
        unsigned lo = 0, hi = ncand;
        while (lo < hi) { unsigned mid = (lo + hi) >> 1; if (sPos[mid] < entry) lo = mid + 1; else hi = mid; }
        if (lo < ncand && sPos[lo] == entry) s_eidx = lo;
    }
    __syncthreads();
    const unsigned eidx = s_eidx;
    if (eidx == JB_U32_NONE) { if (tid == 0) jb_set_error(f.status, JB_ERR_BAD_STREAM); return; }
    for (unsigned j = tid; j < hops; j += JB_FRAME_THREADS) {
        unsigned node = eidx;
        for (int k = 0; k < levels && node != JB_TERM; ++k)
            if ((j >> k) & 1u) node = sLvl[(size_t)k * JB_TILE_BYTES + node];
        if (node == JB_TERM || base + j >= (unsigned)f.nblocks) { jb_set_error(f.status, JB_ERR_BAD_STREAM); continue; }
        f.block_start[(size_t)s * f.nblocks + base + j] = tstart + sPos[node];
    }
}

cudaError_t jb_launch_framing(const JbFrameArgs& f, cudaStream_t s) {
    cudaError_t e;
    jb_frame_prep_kernel<<<1, 1024, 0, s>>>(f);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    size_t smem1 = jb_tile_smem(f.maxblk).total;
    e = cudaFuncSetAttribute(jb_frame_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    if (e != cudaSuccess) return e;
    jb_frame_tiles_kernel<<<f.max_tiles, JB_FRAME_THREADS, smem1, s>>>(f);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    jb_frame_resolve_kernel<<<f.n_planes, JB_FRAME_THREADS, 0, s>>>(f);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    size_t smem3 = (size_t)(1 + JB_EMIT_LEVELS) * JB_TILE_BYTES * 2;
    e = cudaFuncSetAttribute(jb_frame_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3);
    if (e != cudaSuccess) return e;
    jb_frame_emit_kernel<<<f.max_tiles, JB_FRAME_THREADS, smem3, s>>>(f);
    return cudaGetLastError();
}
