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
    unsigned long long* status;       // JB_STATUS_WORDS
    uint8_t* tmp;                     // n_chunks * chunk_cap: each chunk's packed bytes, compacted (chunks longer than JB_SLOT_SMALL)
    uint8_t* tmp_small;               // n_chunks * JB_SLOT_STRIDE: the same for chunks of at most JB_SLOT_SMALL bytes (dense: the
                                      // gather pass then reads ~1 KB strides instead of worst-case-sized ones)
    unsigned chunk_cap;               // bytes reserved per chunk in tmp (multiple of 16)
    unsigned* chunk_len;              // n_chunks: packed bytes of each chunk
    unsigned* chunk_off;              // n_chunks: offset of the chunk inside its scan segment
    unsigned long long* seg_total;    // per scan segment (JB_SCAN_SEG chunks): bytes, then exclusive base
    unsigned* ticket;                 // chunk ticket counter (zeroed)
    int16_t* coeffs_out;              // MODE 1
    const int32_t* coeffs_in;         // MODE 2
};

#define JB_SCAN_SEG 1024            // chunks per scan segment (one CTA of the local scan)

size_t jb_fwd_generic_smem_bytes(int d, bool dft);
// device-wide exclusive scan of chunk lengths + gather of the chunks into the output stream
cudaError_t jb_launch_scan_gather(const JbFwdArgs& a, cudaStream_t s);
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
