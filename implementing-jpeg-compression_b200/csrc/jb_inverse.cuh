// jb_inverse.cuh -- launch arguments of the decompress-direction kernels.
#pragma once
#include "jb_common.cuh"

#define JB_FRAME_THREADS 256
#define JB_INV_GENERIC_THREADS 256

// Workspace of the decoder behind the tables (all offsets 256-byte aligned).
struct JbDecLayout {
    size_t tile_first;   // uint32 [n_planes + 1]   first tile of each stream; [n_planes] = tile count
    size_t block_start;  // uint32 [n_planes * nblocks]  byte offset of every block inside its stream
    size_t tile_uniq;    // uint32 [max_tiles]  exit shared by every entry candidate, or NONE
    size_t tile_entry;   // uint32 [max_tiles]  offset of the first true block start inside the tile
    size_t tile_base;    // uint32 [max_tiles]  ordinal (within the stream) of that block
    size_t tile_hops;    // uint32 [max_tiles]  true blocks that start inside the tile
    size_t tile_ncand;   // uint32 [max_tiles]
    size_t win;          // uint32 [max_tiles * win_n]  (hops << 16 | exit) per entry offset < win_n
    size_t cand_pos;     // uint16 [max_tiles * JB_TILE_BYTES]
    size_t cand_next;    // uint16 [max_tiles * JB_TILE_BYTES]
    size_t total;
    unsigned max_tiles;
    unsigned win_n;
};

static inline JbDecLayout jb_dec_layout(int d, int n_planes, long long nblocks_per_plane, size_t in_bytes,
                                        size_t table_bytes) {
    JbDecLayout L;
    size_t o = jb_align_up(table_bytes, 256);
    L.max_tiles = (unsigned)(in_bytes / JB_TILE_BYTES + (size_t)n_planes + 1);
    L.win_n = (unsigned)jb_max_block_bytes(d * d);
    L.tile_first = o;  o += jb_align_up(((size_t)n_planes + 1) * 4, 256);
    L.block_start = o; o += jb_align_up((size_t)n_planes * (size_t)nblocks_per_plane * 4, 256);
    L.tile_uniq = o;   o += jb_align_up((size_t)L.max_tiles * 4, 256);
    L.tile_entry = o;  o += jb_align_up((size_t)L.max_tiles * 4, 256);
    L.tile_base = o;   o += jb_align_up((size_t)L.max_tiles * 4, 256);
    L.tile_hops = o;   o += jb_align_up((size_t)L.max_tiles * 4, 256);
    L.tile_ncand = o;  o += jb_align_up((size_t)L.max_tiles * 4, 256);
    L.win = o;         o += jb_align_up((size_t)L.max_tiles * L.win_n * 4, 256);
    L.cand_pos = o;    o += jb_align_up((size_t)L.max_tiles * JB_TILE_BYTES * 2, 256);
    L.cand_next = o;   o += jb_align_up((size_t)L.max_tiles * JB_TILE_BYTES * 2, 256);
    L.total = o;
    return L;
}

struct JbFrameArgs {
    const uint8_t* in;
    const unsigned long long* plane_off;
    const unsigned long long* plane_len;
    int n_planes;
    int n;                 // coefficients per block
    int nblocks;           // blocks per plane expected
    int maxblk;
    unsigned max_tiles, win_n;
    unsigned* tile_first;
    unsigned* block_start;
    unsigned* tile_uniq;
    unsigned* tile_entry;
    unsigned* tile_base;
    unsigned* tile_hops;
    unsigned* tile_ncand;
    unsigned* win;
    uint16_t* cand_pos;
    uint16_t* cand_next;
    unsigned long long* status;
};

struct JbInvArgs {
    JbGeom g;
    JbTables t;
    const uint8_t* in;
    const unsigned long long* plane_off;
    const unsigned long long* plane_len;
    const unsigned* block_start;
    int n_planes;
    unsigned n_chunks;
    uint8_t* planes_out;
    size_t plane_stride, row_pitch;
    int16_t* coeffs_out;        // MODE 1
    const int16_t* coeffs_in;   // MODE 2
    unsigned long long* status;
};

cudaError_t jb_launch_framing(const JbFrameArgs& f, cudaStream_t s);
cudaError_t jb_launch_inv_generic(const JbInvArgs& a, int mode, cudaStream_t s);
size_t jb_inv_generic_smem_bytes(int d, bool dft);

bool jb_inv_fast_eligible(const JbGeom& g);
cudaError_t jb_launch_inv_fast(const JbInvArgs& a, int mode, cudaStream_t s);
