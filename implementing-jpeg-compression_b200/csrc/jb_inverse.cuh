// jb_inverse.cuh -- launch arguments of the decompress-direction kernels.
#pragma once
#include <string.h>

#include "jb_common.cuh"

#define JB_FRAME_THREADS 256
#define JB_INV_GENERIC_THREADS 256

// Walk tile (bytes of stream per walking thread): about twelve average blocks, so that a walk which
// starts on a false offset has re-joined the true chain well before its tile ends (a walk that has
// not sends its whole stream to the serial fallback).  Blocks may be longer than a tile: tiles that
// a block spans entirely are simply not on the chain of tiles (jb_frame_reach_kernel).
static inline unsigned jb_frame_tile_bytes(size_t in_bytes, int n_planes, long long nblocks_per_plane) {
    const double blocks = (double)n_planes * (double)nblocks_per_plane;
    const double avg = blocks > 0 ? (double)in_bytes / blocks : 16.0;
#ifndef JB_MIN_TILE
#define JB_MIN_TILE 256
#endif
#ifndef JB_TILE_BLOCKS
#define JB_TILE_BLOCKS 12.0
#endif
    unsigned t = JB_MIN_TILE;
    while (t < 4096 && (double)t < JB_TILE_BLOCKS * avg) t <<= 1;
    return t;
}

// Workspace of the decoder behind the tables (all offsets 256-byte aligned).
struct JbDecLayout {
    size_t ctrl;         // control block (jb_common.cuh): chunk ticket, finished CTAs, status words
    size_t tile_first;   // uint32 [n_planes + 1]   first tile of each stream; [n_planes] = tile count
    size_t fallback;     // uint32 [n_planes]       stream needs the serial walk
    size_t notplain;     // uint32 [n_planes]       some walk of the stream did not leave its tile for the next one
    size_t big_list;     // uint32 [n_planes + 1]   [0] = count, then the streams with more than 4096 tiles
    size_t block_start;  // uint32 [n_planes * nblocks]  byte offset of every block inside its stream
    size_t tile_exit;    // uint32 [max_tiles]  offset where the walk leaves the tile (or invalid)
    size_t tile_entry;   // uint32 [max_tiles]  offset of the first true block start of the tile
    size_t tile_from;    // uint32 [max_tiles]  byte of the tile where the true chain joins the walk (tile_bytes: nowhere)
    size_t tile_npriv;   // uint32 [max_tiles]  true blocks before that point
    size_t tile_hops;    // uint32 [max_tiles]  true blocks that start inside the tile
    size_t tile_base;    // uint32 [max_tiles]  ordinal (within the stream) of the tile's first true block
    size_t vbits;        // uint32 [max_tiles * tile_bytes / 32]  bit b of a tile: its walk saw a block start at byte b
    size_t warp_first;   // uint32 [n_planes + 1]   first warp of each stream in the walk (32 tiles per warp); [n_planes] = warps
    size_t warp_stream;  // uint32 [max_tiles / 32 + n_planes + 1]   stream of every walk warp
    size_t total;
    unsigned max_tiles;
    unsigned tile_bytes;
};

static inline JbDecLayout jb_dec_layout(int d, int n_planes, long long nblocks_per_plane, size_t in_bytes,
                                        size_t table_bytes) {
    JbDecLayout L;
    size_t o = jb_align_up(table_bytes, 256);
    (void)d;
    L.tile_bytes = jb_frame_tile_bytes(in_bytes, n_planes, nblocks_per_plane);
    L.max_tiles = (unsigned)(in_bytes / L.tile_bytes + (size_t)n_planes + 1);
    L.ctrl = o;        o += JB_CTRL_BYTES;
    L.tile_first = o;  o += jb_align_up(((size_t)n_planes + 1) * 4, 256);
    L.fallback = o;    o += jb_align_up((size_t)n_planes * 4, 256);
    L.notplain = o;    o += jb_align_up((size_t)n_planes * 4, 256);
    L.big_list = o;    o += jb_align_up(((size_t)n_planes + 1) * 4, 256);
    L.block_start = o; o += jb_align_up((size_t)n_planes * (size_t)nblocks_per_plane * 4, 256);
    L.tile_exit = o;   o += jb_align_up((size_t)L.max_tiles * 4, 256);
    L.tile_entry = o;  o += jb_align_up((size_t)L.max_tiles * 4, 256);
    L.tile_from = o;   o += jb_align_up((size_t)L.max_tiles * 4, 256);
    L.tile_npriv = o;  o += jb_align_up((size_t)L.max_tiles * 4, 256);
    L.tile_hops = o;   o += jb_align_up((size_t)L.max_tiles * 4, 256);
    L.tile_base = o;   o += jb_align_up((size_t)L.max_tiles * 4, 256);
    L.vbits = o;       o += jb_align_up((size_t)L.max_tiles * (L.tile_bytes / 8), 256);
    L.warp_first = o;  o += jb_align_up(((size_t)n_planes + 1) * 4, 256);
    L.warp_stream = o; o += jb_align_up(((size_t)L.max_tiles / 32 + (size_t)n_planes + 1) * 4, 256);
    L.total = o;
    return L;
}

// Which framing kernels a decompress call takes (decided on the host from the sizes alone):
//   short streams (a batch of images): prep, walk, then one CTA per stream does the rest (jb_frame_stitch_kernel);
//   long streams (a gigapixel plane):  prep, walk, reach, link, scan, emit, serial -- several CTAs per stream.
#define JB_STITCH_CAP 4096u
static inline bool jb_framing_is_chain(unsigned max_tiles, int n_planes) {
    return (size_t)max_tiles > (size_t)n_planes * (JB_STITCH_CAP / 4u);
}

#define JB_SERIAL_WORDS 1024      // words of shared memory the serial fallback walk stages at a time

struct JbFrameArgs {
    const uint8_t* in;
    const unsigned long long* plane_off;
    const unsigned long long* plane_len;
    int n_planes;
    int n;                 // coefficients per block
    int nblocks;           // blocks per plane expected
    int maxblk;
    unsigned max_tiles, tile_bytes;
    int force_serial;      // JB_FLAG_SERIAL_FRAMING: every stream takes the serial walk (test hook)
    int pdl;               // JB_FLAG_PDL
    unsigned stitch_cap;   // set by jb_launch_framing: tiles per stream the one-launch path handles
    size_t in_bytes;       // bytes readable at `in`; a stream that reaches beyond is malformed
    unsigned* tile_first;
    unsigned* fallback;
    unsigned* notplain;
    unsigned* big_list;
    unsigned* block_start;
    unsigned* tile_exit;
    unsigned* tile_entry;
    unsigned* tile_from;
    unsigned* tile_npriv;
    unsigned* tile_hops;
    unsigned* tile_base;
    uint32_t* vbits;
    unsigned* warp_first;
    unsigned* warp_stream;
    unsigned long long* status;
};

struct JbInvArgs {
    JbGeom g;
    JbTables t;
    const uint8_t* in;
    size_t in_bytes;            // bytes readable at `in`
    const unsigned long long* plane_off;
    const unsigned long long* plane_len;
    const unsigned* block_start;
    int n_planes;
    unsigned n_chunks;
    uint8_t* planes_out;
    size_t plane_stride, row_pitch;
    int16_t* coeffs_out;        // MODE 1
    const int16_t* coeffs_in;   // MODE 2
    unsigned* ticket;           // chunks are claimed from this counter (zero at launch); nullptr: dealt round-robin
    unsigned* done;             // finished CTAs of the fused inverse kernel (zero at launch)
    unsigned long long* status; // where the kernels record errors: the workspace's control block, or (stage entry
                                // point without framing) the caller's block itself
    unsigned long long* status_out;   // the caller's status words when `status` is the control block, else nullptr
};

// publishes the control block's status words to the caller and leaves the block clean (end of a decompress call)
__device__ __forceinline__ void jb_dec_publish_and_clean(const JbInvArgs& a) {
    #pragma unroll
    for (int k = 0; k < JB_STATUS_WORDS; ++k) {
        a.status_out[k] = __ldcg(a.status + k);
        a.status[k] = k == 1 ? ~0ull : 0ull;
    }
    if (a.ticket) *a.ticket = 0u;
    if (a.done) *a.done = 0u;
}
cudaError_t jb_launch_dec_finish(const JbInvArgs& a, cudaStream_t s);

// End of a decompress call inside a persistent transform kernel: the CTA that finishes last hands the status words
// to the caller and leaves the control block clean -- one atomic per CTA.  Every thread of the CTA must call.
__device__ __forceinline__ void jb_dec_epilogue(const JbInvArgs& a) {
    if (a.status_out == nullptr) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(a.done, 1u) == gridDim.x - 1u) {
            __threadfence();
            jb_dec_publish_and_clean(a);
        }
    }
}

cudaError_t jb_launch_framing(const JbFrameArgs& f, cudaStream_t s);
cudaError_t jb_launch_inv_generic(const JbInvArgs& a, int mode, cudaStream_t s);
size_t jb_inv_generic_smem_bytes(int d, bool dft);

// large-block path: DCT, dct_size 16 / 24 / 32 (jb_inverse_large.cu); ends the call itself like the 8x8 kernel
bool jb_inv_large_eligible(const JbGeom& g);
cudaError_t jb_launch_inv_large(const JbInvArgs& a, int mode, cudaStream_t s);

bool jb_inv_fast_eligible(const JbGeom& g);
cudaError_t jb_launch_inv_fast(const JbInvArgs& a, int mode, cudaStream_t s);
