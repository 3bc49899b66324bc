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
    unsigned long long* seg1;         // [2][n_seg1] bytes per JB_SEG1 chunks, accumulated by the transform kernel
    unsigned long long* seg2;         // [2][n_seg2] bytes per JB_SEG2 chunks, likewise
    unsigned n_seg1, n_seg2;
    int16_t* coeffs_out;              // MODE 1
    const int32_t* coeffs_in;         // MODE 2
};

// Stream offsets without a scan pass: the transform kernels add every chunk's byte count to two levels of
// segment totals (atomics); a gather CTA then sums at most n_chunks / JB_SEG2 + JB_SEG2 / JB_SEG1 + JB_SEG1 values.
#define JB_SEG1 64                  // chunks per first-level segment (a multiple of the chunks per gather CTA)
#define JB_SEG2 4096                // chunks per second-level segment (a multiple of JB_SEG1)

// Control block of the compress workspace (256 bytes behind the tables): two sets of {chunk ticket, status
// words, segment totals}, used alternately.
//   uint32 [0] parity P: the set this call uses (read by the transform kernel; stable while it runs)
//   uint32 [1] the same P, left by the transform kernel for the call's last kernel
//   uint32 [2 + P] chunk ticket of set P
//   uint64 status[P][JB_STATUS_WORDS] at byte 64 + 32 P
// A call works on set P, which is clean when it starts (ticket 0, status {0, ~0, 0, 0}, totals 0).  Its last
// kernel (gather, or jb_finish_kernel) copies status[P] to the caller's block, cleans set P ^ 1 -- last used by
// the call before, so nobody reads it now -- and flips [0].  No memset node, no "last CTA" counter: a call is the
// transform kernel and the gather.  jb_init_ctrl_kernel (calls without JB_FLAG_REUSE_TABLES) cleans both sets.
#define JB_CTRL_BYTES 256
#define JB_CTRL_STATUS_OFF 64

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
    a.chunk_len[chunk] = total;
    atomicAdd(a.seg1 + (size_t)P * a.n_seg1 + chunk / JB_SEG1, (unsigned long long)total);
    atomicAdd(a.seg2 + (size_t)P * a.n_seg2 + chunk / JB_SEG2, (unsigned long long)total);
}

size_t jb_fwd_generic_smem_bytes(int d, bool dft);
// gather of the chunks into the output stream (computes every chunk's offset from the segment totals), then
// publication of the status words and clean-up of the control block
cudaError_t jb_launch_gather(const JbFwdArgs& a, cudaStream_t s);
// the same publication / clean-up for calls that end without a gather pass (stage entry point: coefficients only)
cudaError_t jb_launch_finish(const JbFwdArgs& a, cudaStream_t s);
// cleans both sets of the control block and the segment totals (calls without JB_FLAG_REUSE_TABLES)
cudaError_t jb_launch_init_ctrl(unsigned* ctrl, unsigned long long* seg, size_t n_seg_words, cudaStream_t s);
cudaError_t jb_launch_fwd_generic(const JbFwdArgs& a, int mode, cudaStream_t s);

// large-block path: dct_size a multiple of 4, source tile in shared memory (jb_forward_mid.cu)
bool jb_fwd_mid_eligible(const JbGeom& g);
cudaError_t jb_launch_fwd_mid(const JbFwdArgs& a, int mode, cudaStream_t s);

// specialised path: dct_size 8, block_size 4 (jb_forward_fast.cu)
bool jb_fwd_fast_eligible(const JbGeom& g);
cudaError_t jb_launch_fwd_fast(const JbFwdArgs& a, int mode, cudaStream_t s);

// where a chunk's packed bytes are staged between the fused kernel and the gather pass
__device__ __forceinline__ uint8_t* jb_chunk_slot(const JbFwdArgs& a, unsigned chunk, unsigned total_bytes) {
    return total_bytes <= JB_SLOT_SMALL ? a.tmp_small + (size_t)chunk * JB_SLOT_STRIDE : a.tmp + (size_t)chunk * a.chunk_cap;
}
