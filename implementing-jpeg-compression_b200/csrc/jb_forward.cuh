// jb_forward.cuh -- launch arguments of the compress-direction kernels.
#pragma once
#include "jb_common.cuh"

#define JB_GENERIC_THREADS 256

#define JB_SLOT_SMALL 1024u        // bytes of a chunk that still go to the dense slot array
#define JB_SLOT_STRIDE 1056u       // its stride: 1 KB + the 32 bytes the gather kernel reads ahead

struct JbFwdArgs {
    JbGeom g;
    JbTables t;
    const uint8_t* planes;            // MODE 0/1 input
    size_t plane_stride, row_pitch;
    int n_planes;
    unsigned n_chunks;                // n_planes * g.cpp
    uint8_t* out;                     // MODE 0/2 output: concatenated streams
    unsigned long long out_cap;
    unsigned long long* plane_off;    // n_planes + 1
    unsigned* ctrl;                   // control block of the workspace (below): parity, tickets, status words
    unsigned long long* status_out;   // the caller's status words, written once by the call's last kernel
    uint8_t* tmp;                     // n_chunks * chunk_cap: each chunk's packed bytes, compacted (chunks longer than JB_SLOT_SMALL)
    uint8_t* tmp_small;               // n_chunks * JB_SLOT_STRIDE: the same for chunks of at most JB_SLOT_SMALL bytes (dense: the
                                      // gather pass then reads ~1 KB strides instead of worst-case-sized ones)
    unsigned chunk_cap;               // bytes reserved per chunk in tmp (multiple of 16)
    unsigned* chunk_len;              // n_chunks: packed bytes of each chunk
    unsigned* chunk_off;              // n_chunks: offset of the chunk inside its scan segment
    unsigned long long* seg_total;    // per scan segment (JB_SCAN_SEG chunks): bytes, then exclusive base
    int slots_always_big;             // every chunk was staged in `tmp` (large-block kernel), none in `tmp_small`
    int16_t* coeffs_out;              // MODE 1
    const int32_t* coeffs_in;         // MODE 2
};

#define JB_SCAN_SEG 1024            // chunks per scan segment (one CTA of the scan kernel)

// Control block of the compress workspace (256 bytes behind the tables): two sets of {chunk ticket, status
// words}, used alternately.
//   uint32 [0] parity P: the set this call uses (read by the transform kernel; stable while it runs)
//   uint32 [1] the same P, left by the transform kernel for the kernels behind it
//   uint32 [2 + P] chunk ticket of set P
//   uint32 [4] finished CTAs of the scan kernel (reset by its last CTA)
//   uint64 status[P][JB_STATUS_WORDS] at byte 64 + 32 P
// A call works on set P, which is clean when it starts (ticket 0, status {0, ~0, 0, 0}).  Its last kernel (gather,
// or jb_finish_kernel) copies status[P] to the caller's block, cleans set P ^ 1 -- last used by the call before, so
// nobody reads it now -- and flips [0].  No memset / init node: a call is the transform kernel, one scan kernel and
// the gather.  jb_init_ctrl_kernel (calls without JB_FLAG_REUSE_TABLES) cleans both sets.
__device__ __forceinline__ int jb_ctrl_parity(const JbFwdArgs& a) { return (int)(__ldcg(a.ctrl) & 1u); }
__device__ __forceinline__ unsigned* jb_ctrl_ticket(const JbFwdArgs& a, int P) { return a.ctrl + 2 + P; }
__device__ __forceinline__ unsigned long long* jb_ctrl_status(const JbFwdArgs& a, int P) {
    return (unsigned long long*)(a.ctrl + JB_CTRL_STATUS_OFF / 4) + JB_STATUS_WORDS * P;
}
// the thread that drew ticket 0 hands the parity on to the call's last kernel
__device__ __forceinline__ void jb_ctrl_note_first(const JbFwdArgs& a, int P, unsigned ticket) {
    if (ticket == 0u) a.ctrl[1] = (unsigned)P;
}

__device__ __forceinline__ void jb_record_chunk_len(const JbFwdArgs& a, int P, unsigned chunk, unsigned total) {
    (void)P;
    a.chunk_len[chunk] = total;
}

size_t jb_fwd_generic_smem_bytes(int d, bool dft);
// device-wide exclusive scan of the chunk lengths (one launch) + gather of the chunks into the output stream, which
// also publishes the status words and cleans the control block
cudaError_t jb_launch_gather(const JbFwdArgs& a, cudaStream_t s);
// the same publication / clean-up for calls that end without a gather pass (stage entry point: coefficients only)
cudaError_t jb_launch_finish(const JbFwdArgs& a, cudaStream_t s);

cudaError_t jb_launch_fwd_generic(const JbFwdArgs& a, int mode, cudaStream_t s);

// large-block path: DCT, dct_size 16 / 24 / 32, tile rows of at most 128 bytes (jb_forward_large.cu)
bool jb_fwd_large_eligible(const JbGeom& g);
cudaError_t jb_launch_fwd_large(const JbFwdArgs& a, int mode, cudaStream_t s);

// specialised path: dct_size 8, block_size 4 (jb_forward_fast.cu)
bool jb_fwd_fast_eligible(const JbGeom& g);
cudaError_t jb_launch_fwd_fast(const JbFwdArgs& a, int mode, cudaStream_t s);

// where a chunk's packed bytes are staged between the fused kernel and the gather pass
__device__ __forceinline__ uint8_t* jb_chunk_slot(const JbFwdArgs& a, unsigned chunk, unsigned total_bytes) {
    return total_bytes <= JB_SLOT_SMALL ? a.tmp_small + (size_t)chunk * JB_SLOT_STRIDE : a.tmp + (size_t)chunk * a.chunk_cap;
}
