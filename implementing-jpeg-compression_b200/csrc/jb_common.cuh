// jb_common.cuh -- shared definitions of the libjpegb200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/jpegb200.h"
#include "jb_bits.h"

#define JB_CHUNK 32            // blocks per chunk: one RLE/entropy thread per block, one warp per chunk
#define JB_TILE_BYTES 4096     // stream bytes per framing tile (decoder)
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

// ---- decoupled look-back over chunk byte lengths ----------------------------------------
// desc[c] = flag << 62 | value; flag 0 = not ready, 1 = aggregate (this chunk's bytes),
// 2 = inclusive prefix (bytes of chunks 0..c).
#define JB_DESC_AGG (1ull << 62)
#define JB_DESC_PFX (2ull << 62)
#define JB_DESC_VAL (~(3ull << 62))

__device__ __forceinline__ unsigned long long jb_ld_relaxed(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void jb_st_relaxed(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// Called by a full warp; `mine` = this chunk's byte count (uniform).  Returns the
// exclusive prefix (bytes of all earlier chunks).  Chunks must be claimed in
// increasing order (atomic ticket) so that every predecessor is already running.
// A bounded spin turns a lost descriptor into an error instead of a hang.
__device__ __forceinline__ unsigned long long jb_lookback_exclusive(unsigned long long* desc, unsigned chunk,
                                                                    unsigned long long mine, int lane,
                                                                    unsigned long long* status) {
    if (chunk == 0) {
        if (lane == 0) jb_st_relaxed(desc, JB_DESC_PFX | mine);
        return 0ull;
    }
    if (lane == 0) jb_st_relaxed(desc + chunk, JB_DESC_AGG | mine);
    unsigned long long excl = 0ull;
    long long look = (long long)chunk - 1;     // highest predecessor not yet accounted for
    int spins = 0;
    while (true) {
        long long idx = look - lane;
        unsigned long long v = (idx >= 0) ? jb_ld_relaxed(desc + idx) : JB_DESC_PFX;   // virtual prefix 0 before chunk 0
        unsigned flag = (unsigned)(v >> 62);
        unsigned not_ready = __ballot_sync(0xffffffffu, flag == 0u);
        unsigned has_pfx = __ballot_sync(0xffffffffu, flag == 2u);
        // usable window: lanes below the first not-ready lane; stop at the first prefix
        int first_nr = not_ready ? __ffs(not_ready) - 1 : 32;
        int first_pf = has_pfx ? __ffs(has_pfx) - 1 : 32;
        int take = first_pf < first_nr ? first_pf + 1 : first_nr;   // number of lanes to add
        unsigned long long contrib = (lane < take) ? (v & JB_DESC_VAL) : 0ull;
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        excl += contrib;
        if (first_pf < first_nr) break;          // reached an inclusive prefix
        look -= take;
        if (take == 0) {
            if (++spins > (1 << 22)) {           // ~seconds: give up loudly, never hang the GPU
                if (lane == 0) jb_set_error(status, JB_ERR_CUDA);
                break;
            }
            __nanosleep(64);
        } else {
            spins = 0;
        }
    }
    if (lane == 0) jb_st_relaxed(desc + chunk, JB_DESC_PFX | (excl + mine));
    return excl;
}

// ---- small helpers -------------------------------------------------------------------------
__device__ __forceinline__ int jb_min(int a, int b) { return a < b ? a : b; }

// round-half-even to integer for |x| < 2^22 without a conversion instruction
__device__ __forceinline__ float jb_rint_magic(float x) {
    return (x + 12582912.0f) - 12582912.0f;
}

// Host-side API internals shared between translation units.
int jb_make_geom(const jb_params* p, JbGeom* g);
cudaError_t jb_launch_build_tables(const JbGeom& g, const JbTables& t, cudaStream_t s);
