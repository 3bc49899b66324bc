// jb_common.cuh -- shared definitions of the libjpegb200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/jpegb200.h"
#include "jb_bits.h"

#define JB_CHUNK 32            // blocks per chunk: one RLE/entropy thread per block, one warp per chunk
#define JB_CHUNK_LARGE 8       // ... for dct_size >= 16, where a block is hundreds of coefficients and a chunk of 32
                               // would be megabytes of pixels per scheduling unit (a 4K frame: 54 units for 148 SMs)
__host__ __device__ inline int jb_chunk_blocks(int d) { return d >= 16 ? JB_CHUNK_LARGE : JB_CHUNK; }
// Large blocks, small calls: a 4K frame of 24 x 24 blocks is 216 chunks of 8 for some 1500 resident warps, each of which
// then works through its 8 blocks one after the other.  Calls of up to JB_SMALL_CALL_BLOCKS blocks in total are dealt in
// chunks of JB_CHUNK_LARGE_SMALL blocks instead (2: the decoder's pixel rows of a chunk still start on 16-byte
// boundaries for even block columns).  Both directions, workspace sizes included, derive the chunk from this one rule.
#define JB_CHUNK_LARGE_SMALL 2
#define JB_SMALL_CALL_BLOCKS 16384ll
__host__ __device__ inline int jb_call_chunk_blocks(int d, long long total_blocks) {
    if (d >= 16 && total_blocks <= JB_SMALL_CALL_BLOCKS) return JB_CHUNK_LARGE_SMALL;
    return jb_chunk_blocks(d);
}

#define JB_U32_NONE 0xFFFFFFFFu
#define JB_U16_NONE 0xFFFFu

// Geometry of one plane, host + device (reference: run_length_encoding.py:80-88).
struct JbGeom {
    int H, W;          // source plane
    int bs, d, n;      // block_size, dct_size, d*d
    int H1, W1;        // subsampled
    int H2, W2;        // padded to multiples of d
    int vb, hb;        // blocks down / across
    int nblocks;       // vb * hb
    int maxblk;        // worst-case bytes per block
    int cpp;           // chunks per plane
    int chunk;         // blocks per chunk: jb_chunk_blocks(d)
    int transform, qmode, qparam, flags;
};

// Device tables, built per call by jb_build_tables_kernel into the workspace.
// Natural index = u * d + v (u = vertical frequency / row).
struct JbTables {
    float*    fA;      // [d*d] forward matrix, fp32: DCT C[k][m] = cos(pi/N (m+.5) k); DFT cos(2 pi k m / N)
    float*    fB;      // [d*d] DFT only: sin(2 pi k m / N)
    double*   fA64;    // same in fp64 (near-tie re-evaluation)
    double*   fB64;
    float*    qmult;   // [d*d] v = Ysum * qmult  (Ysum = transform of the integer box sums)
    float*    qtol;    // [d*d] |frac(v) - .5| below this => re-evaluate in fp64 (negative: never)
    double*   qrecip;  // [d*d] fp64: 1.0/q (qtable), divisor (divide), 1 (none), 0 (discarded)
    float*    iA;      // [d*d] inverse matrix fp32: DCT B[m][k] = C[k][m]/|c_k|^2 ; DFT cos(..)/N
    float*    iB;      // [d*d] DFT only: sin(..)/N
    float*    dqmult;  // [d*d] dequantiser: q (qtable), divisor (divide), else 1
    uint16_t* zz;      // [d*d] natural index -> zigzag position
    uint16_t* izz;     // [d*d] zigzag position -> natural index
};

__host__ __device__ inline size_t jb_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// byte offsets of the tables inside the workspace (same on host and device)
struct JbTableLayout {
    size_t fA, fB, fA64, fB64, qmult, qtol, qrecip, iA, iB, dqmult, zz, izz, total;
};

static inline JbTableLayout jb_table_layout(int d) {
    JbTableLayout L;
    size_t n = (size_t)d * d, o = 0;
    L.fA64 = o;   o += jb_align_up(n * 8, 256);
    L.fB64 = o;   o += jb_align_up(n * 8, 256);
    L.qrecip = o; o += jb_align_up(n * 8, 256);
    L.fA = o;     o += jb_align_up(n * 4, 256);
    L.fB = o;     o += jb_align_up(n * 4, 256);
    L.qmult = o;  o += jb_align_up(n * 4, 256);
    L.qtol = o;   o += jb_align_up(n * 4, 256);
    L.iA = o;     o += jb_align_up(n * 4, 256);
    L.iB = o;     o += jb_align_up(n * 4, 256);
    L.dqmult = o; o += jb_align_up(n * 4, 256);
    L.zz = o;     o += jb_align_up(n * 2, 256);
    L.izz = o;    o += jb_align_up(n * 2, 256);
    L.total = o;
    return L;
}

static inline JbTables jb_tables_at(void* ws, int d) {
    JbTableLayout L = jb_table_layout(d);
    char* b = (char*)ws;
    JbTables t;
    t.fA = (float*)(b + L.fA);       t.fB = (float*)(b + L.fB);
    t.fA64 = (double*)(b + L.fA64);  t.fB64 = (double*)(b + L.fB64);
    t.qmult = (float*)(b + L.qmult); t.qtol = (float*)(b + L.qtol);
    t.qrecip = (double*)(b + L.qrecip);
    t.iA = (float*)(b + L.iA);       t.iB = (float*)(b + L.iB);
    t.dqmult = (float*)(b + L.dqmult);
    t.zz = (uint16_t*)(b + L.zz);    t.izz = (uint16_t*)(b + L.izz);
    return t;
}

// ---- status words ------------------------------------------------------------------
// status[0] = positive error code (atomicMax), status[1] = packed first bad RLE code
// (atomicMin): block:30 | pos:10 | run:4 | (amp + 2^19):20
#define JB_BADCODE_AMP_BIAS (1 << 19)

__device__ __forceinline__ void jb_set_error(unsigned long long* status, int neg_code) {
    atomicMax(status, (unsigned long long)(-neg_code));
}

__device__ __forceinline__ void jb_report_bad_code(unsigned long long* status, unsigned long long global_block,
                                                   int pos, int run, long long amp) {
    long long biased = amp + JB_BADCODE_AMP_BIAS;
    if (biased < 0) biased = 0;
    if (biased > 0xFFFFF) biased = 0xFFFFF;
    unsigned long long w = ((global_block & 0x3FFFFFFFull) << 34) | ((unsigned long long)(pos & 1023) << 24)
                         | ((unsigned long long)(run & 15) << 20) | (unsigned long long)biased;
    atomicMin(status + 1, w);
    jb_set_error(status, JB_ERR_BAD_RLE_CODE);
}

// ---- small helpers -------------------------------------------------------------------------
__device__ __forceinline__ int jb_min(int a, int b) { return a < b ? a : b; }

// round-half-even to integer for |x| < 2^22 without a conversion instruction
__device__ __forceinline__ float jb_rint_magic(float x) {
    return (x + 12582912.0f) - 12582912.0f;
}

__device__ __forceinline__ unsigned jb_warp_excl_scan(unsigned v, int lane, unsigned* total) {
    unsigned incl = v;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    *total = __shfl_sync(0xffffffffu, incl, 31);
    return incl - v;
}

// exclusive scan across a block of up to 1024 threads; every thread must call
__device__ __forceinline__ unsigned jb_block_excl_scan(unsigned v, unsigned* s_warp /*[33]*/, unsigned* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    unsigned wtot;
    unsigned ex = jb_warp_excl_scan(v, lane, &wtot);
    __syncthreads();                       // protect s_warp from the previous use
    if (lane == 0) s_warp[warp] = wtot;
    __syncthreads();
    if (warp == 0) {
        unsigned w = lane < nw ? s_warp[lane] : 0u, t;
        unsigned wex = jb_warp_excl_scan(w, lane, &t);
        s_warp[lane] = wex;
        if (lane == 0) s_warp[32] = t;
    }
    __syncthreads();
    *total = s_warp[32];
    return ex + s_warp[warp];
}

// Control block of a workspace (256 bytes behind the tables).  Compress: see jb_forward.cuh.  Decompress:
//   uint32 [0] chunk ticket of the fused inverse kernel, uint32 [1] its finished CTAs,
//   uint64 status[JB_STATUS_WORDS] at byte 64: the kernels of a call record errors HERE (atomics);
// all of it is clean (zero, status[1] all ones) when a call starts, and the call's last kernel copies the status
// words to the caller's block and leaves it clean again.  A call without JB_FLAG_REUSE_TABLES cleans it first.
#define JB_CTRL_BYTES 256
#define JB_CTRL_STATUS_OFF 64
cudaError_t jb_launch_init_ctrl(unsigned* ctrl, unsigned long long* seg, size_t n_seg_words, cudaStream_t s);

// ---- programmatic dependent launch (JB_FLAG_PDL) ----------------------------------------------------------
// The kernels of a call, and the calls that follow each other on a stream, are short enough for launch latency
// and kernel prologues to matter (128 images per rank: ~170 us kernels).  Launched with the programmatic-stream-
// serialization attribute, a kernel's CTAs become resident while the tail of its predecessor still runs; it does
// its prologue (shared-memory setup, constant tables) and then waits in jb_pdl_wait() until the predecessor has
// completed and its writes are visible.  Every kernel of the hot chain triggers its dependents at once and waits
// before the first access to anything a predecessor may write.  Without the attribute both are no-ops.
__device__ __forceinline__ void jb_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void jb_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Every kernel launch of the library goes through JB_LAUNCH or jb_launch_ex and is counted (process-wide, relaxed
// atomic): jb_debug_launch_count() lets the benchmark report how many kernels a call really launches.
void jb_note_launches(unsigned n);
#define JB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    do { kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); jb_note_launches(1u); } while (0)

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t jb_launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                       bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    jb_note_launches(1u);
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// Host-side API internals shared between translation units.
int jb_make_geom(const jb_params* p, JbGeom* g);
cudaError_t jb_launch_build_tables(const JbGeom& g, const JbTables& t, cudaStream_t s);
// Measurement hook (jb_debug_kernel_events): records event `which` (0/1 around the fused forward kernel, 2/3
// around the fused inverse kernel) on the stream if the caller armed it; a no-op otherwise.
void jb_prof_mark(int which, cudaStream_t s);
