// jb_forward_large.cu -- compress direction for large blocks: DCT, dct_size 16 / 24 / 32, source tile rows of at
// most 128 bytes (BASELINE.json config 3: --block_size 5 --dct_size 24 --quantization divide --qdivisor 1000).
//
// Work unit: a chunk of g.chunk = 8 (2 in small calls: jb_call_chunk_blocks) consecutive blocks of one plane, claimed from the ticket counter by a
// WARP of a persistent grid (one CTA of FL_WARPS warps per SM); the warp takes its blocks through every stage on
// its own, so nothing in the kernel is CTA-wide after the table preload:
//   * the (d bs) x (d bs) source tile of a block arrives in the warp's shared-memory slot by TMA (one elected lane,
//     box 144 bytes x d bs rows from the 16-byte boundary at or before the block, mbarrier completion); the load of the NEXT block is issued as soon as the box sums
//     have consumed the tile, so it overlaps the arithmetic.  Edge blocks (padding.py:8-12, dct_padding.py:8-9
//     replicate, TMA would zero-fill) and unaligned planes are filled by the warp with clamped loads;
//   * box sums (subsampling.py:9-11 without the division): lane j sums the bs x bs bytes of sample (i, j) with
//     __dp4a under byte masks, row i by row i; the sums of columns j and d-1-j meet in one lane, which stores
//     Xe = S[j] + S[d-1-j] and Xo = S[j] - S[d-1-j] (exact integers);
//   * C.X.C^T (transforms.py:46-58) as two half-length contractions: the DCT-II matrix is (anti)symmetric,
//     C[k][d-1-j] = (-1)^k C[k][j], so even frequencies see only Xe and odd ones only Xo -- d/2 instead of d
//     multiply-adds per output.  Lane (rg, kg) owns a (d/8) x (d/4) register tile of outputs; operands come from
//     shared memory as 128-bit rows (bank-conflict free for d = 24); the column pass gets its even / odd sums of
//     the row-pass results by one lane exchange (__shfl_xor 28) and a trip through shared memory.
//     fp32 on exact integers; a coefficient within the proven fp32 error bound of a rounding tie is re-evaluated in
//     fp64 in the reference's order (jb_refine.cuh) -- the bound of jb_tables.cu covers the half-length sums too
//     (|Xe|, |Xo| <= 2 M, half as many terms);
//   * quantise, zigzag (int16 row in shared memory), non-zero bitmap by warp ballots;
//   * run-length code + bit packing by the whole warp (util.py:146-160, 203-221): lane w owns the 32 zigzag positions
//     of bitmap word w, a warp scan of the code lengths gives every lane its bit offset, the codes are OR-ed into the
//     block's bytes in shared memory; the bytes go straight to the chunk's slot (blocks of a chunk are produced
//     in order by one warp, so the slot is compact), and the gather pass of jb_forward.cu places the chunks.
// No tensor-core variant: after the half-length trick the contraction is ~1 MAC per source byte; ncu
// (profiles/) shows the FMA pipe far from saturated -- SURVEY.md section 7.2.
#include <cuda.h>
#include <type_traits>
#include <string.h>

#include "jb_common.cuh"
#include "jb_fast_common.cuh"
#include "jb_forward.cuh"
#include "jb_refine.cuh"

#define FL_WARPS 10                    // at most; fewer when the per-warp buffers of a geometry do not fit (fl_layout)
#define FL_BLOCK_BUF_WORDS 96          // 384 bytes of packed output per block in shared memory; longer blocks: serial path
#define FL_TILE_ROW 144                // bytes per tile row in shared memory = TMA box width: a tile row of up to 128 bytes
                                       // plus up to 15 bytes in front of it -- the box has to start on a 16-byte
                                       // boundary of the plane row (block columns are d bs bytes apart: 120 for config 3)
#define FL_BIG_CAP 8                   // amplitudes beyond the 15-bit size field remembered per block (for the error report)
#define FL_RQ_CAP 60                   // near-tie coefficients of a block queued for the warp's float64 re-evaluation

struct FlLayout {
    int h;                  // d / 2
    int xs;                 // floats per row of Xe / Xo (and of the half matrix)
    int ts;                 // floats per row of Te / To
    size_t ch, qm, thr, zz;                         // CTA-wide tables
    size_t warp0, tile, xe, te, crow, obuf, big, rq, bar, warp_bytes, total;
    int warps;              // warps per CTA that fit the 227 KB of one SM
};

__host__ __device__ inline FlLayout fl_layout(int d, int side) {
    FlLayout L;
    const int n = d * d;
    L.h = d / 2;
    L.xs = L.h;                                     // 12 floats = 48 bytes: 16-byte aligned rows, conflict-free for d = 24
    L.ts = d;
    size_t o = 0;
    L.ch = o;   o += jb_align_up((size_t)d * L.xs * 4, 16);       // CH[row(k)][j < h]
    L.qm = o;   o += (size_t)n * 4;
    L.thr = o;  o += (size_t)n * 4;
    L.zz = o;   o += jb_align_up((size_t)n * 2, 16);
    L.warp0 = jb_align_up(o, 128);
    size_t w = 0;
    L.tile = w; w += jb_align_up((size_t)side * FL_TILE_ROW, 128);
    L.xe = w;   w += jb_align_up((size_t)2 * d * L.xs * 4, 16);   // Xe rows, then Xo rows
    L.te = w;   w += jb_align_up((size_t)2 * L.h * L.ts * 4, 16); // Te rows (j < h), then To rows
    // Two buffers live inside others whose contents are dead by then: the quantised coefficients (written once the
    // whole warp has left the column pass) in Te / To, the packed bytes of the block (built after the float64
    // re-evaluation, the last reader of the box sums) in Xe / Xo -- the kilobyte and a half per warp is what puts a tenth
    // warp on the SM for config 3.
    L.crow = L.te;
    L.obuf = L.xe;
    L.big = w;  w += (size_t)(1 + 2 * FL_BIG_CAP) * 4;
    L.rq = w;   w += 4 + (size_t)FL_RQ_CAP * 2 + 4;
    w = jb_align_up(w, 16);
    L.bar = w;  w += 16;
    L.warp_bytes = jb_align_up(w, 128);
    const size_t budget = 227 * 1024 - FL_WARPS * 32 * 4;       // (the kernel's static shared memory)
    L.warps = (int)((budget - L.warp0) / L.warp_bytes);
    if (L.warps > FL_WARPS) L.warps = FL_WARPS;
    L.total = L.warp0 + (size_t)(L.warps > 0 ? L.warps : 1) * L.warp_bytes;
    return L;
}

bool jb_fwd_large_eligible(const JbGeom& g) {
    if (g.transform != JB_TRANSFORM_DCT) return false;
    if (g.d != 16 && g.d != 24 && g.d != 32) return false;
    const int side = g.d * g.bs;
    return side <= 128 && fl_layout(g.d, side).warps >= 4;
}

struct FlKernelArgs {
    JbFwdArgs a;
    int use_tma;        // tensor map valid (aligned planes at least one tile large)
};

// fp64 re-evaluation (jb_refine.cuh) from the even / odd sums: X[i][j] = (Xe + Xo) / 2, X[i][d-1-j] = (Xe - Xo) / 2
struct FlBoxSums {
    const float* xe;
    const float* xo;
    int d, h, xs;
    __device__ __forceinline__ float operator[](int idx) const {
        const int i = idx / d, j = idx - i * d;
        const int jj = j < h ? j : d - 1 - j;
        const float e = xe[i * xs + jj], o = xo[i * xs + jj];
        return (j < h ? e + o : e - o) * 0.5f;
    }
};

template <int D>
__device__ __noinline__ double fl_refine(const float* xe, int uv, int bs_qmode, const double* A64, double recip) {
    FlBoxSums X = {xe, xe + D * (D / 2), D, D / 2, D / 2};
    return jb_refine_f64(X, uv / D, uv % D, D, bs_qmode >> 8, JB_TRANSFORM_DCT, bs_qmode & 255, A64, A64, recip);
}

// serial packing of one block straight to global memory (blocks longer than the shared-memory buffer)
__device__ __noinline__ unsigned fl_pack_serial(const int16_t* c, const uint32_t* mask, int mask_words, uint8_t* dst) {
    JbByteWriter bw;
    bw.init(dst);
    int prev = -1;
    for (int wi = 0; wi < mask_words; ++wi) {
        uint32_t m = mask[wi];
        while (m) {
            const int p = wi * 32 + __ffs((int)m) - 1;
            m &= m - 1;
            const int amp = c[p];
            jb_put_coefficient(bw, p - prev - 1, amp);          // (amplitudes that do not fit were reported already)
            prev = p;
        }
    }
    bw.put(0u, 8);
    return bw.finish();
}

// BS: block_size as a compile-time constant (5: config 3) for fully unrolled box sums, or 0: any block_size at run time.
template <int D, int MODE, int BS>
__global__ void __launch_bounds__(FL_WARPS * 32, 1)
jb_fwd_large_kernel(const __grid_constant__ CUtensorMap tmap, const FlKernelArgs ka) {
    extern __shared__ __align__(128) unsigned char smem[];
    const JbFwdArgs& a = ka.a;
    const JbGeom& g = a.g;
    constexpr int n = D * D, H = D / 2, RT = D / 8, KT = D / 4, NW = n / 32;
    const int bs = BS ? BS : g.bs, side = D * bs;
    const FlLayout L = fl_layout(D, side);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* sCH = (float*)(smem + L.ch);
    float* sQm = (float*)(smem + L.qm);
    float* sThr = (float*)(smem + L.thr);
    uint16_t* sZz = (uint16_t*)(smem + L.zz);
    unsigned char* wbase = smem + L.warp0 + (size_t)warp * L.warp_bytes;
    uint8_t* tile = wbase + L.tile;
    float* sXe = (float*)(wbase + L.xe);
    float* sXo = sXe + D * L.xs;
    float* sTe = (float*)(wbase + L.te);
    float* sTo = sTe + H * L.ts;
    int16_t* crow = (int16_t*)(wbase + L.crow);
    uint32_t* obuf = (uint32_t*)(wbase + L.obuf);
    unsigned long long* bar = (unsigned long long*)(wbase + L.bar);
    int* big = (int*)(wbase + L.big);                    // [0] = count, then (zigzag position, amplitude) pairs
    int* rq_count = (int*)(wbase + L.rq);                // near-tie coefficients of the current block: count, then natural indices
    uint16_t* rq = (uint16_t*)(wbase + L.rq + 4);

    jb_pdl_trigger();
    const bool tables_const = (g.flags & JB_FLAG_REUSE_TABLES) != 0;
    if (!tables_const) jb_pdl_wait();
    // CH[row][j]: row = parity * H + m holds frequency k = 2 m + parity, columns j < H (the other half follows by symmetry)
    for (int idx = threadIdx.x; idx < D * H; idx += blockDim.x) {
        const int row = idx / H, j = idx - row * H;
        const int k = 2 * (row % H) + row / H;
        sCH[row * L.xs + j] = a.t.fA[k * D + j];
    }
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        sQm[idx] = a.t.qmult[idx];
        // near a tie  <=>  |frac - .5| < tol + 2.4e-7 |val|  <=>  |val - rint(val)| + 2.4e-7 |val| > .5 - tol
        const float tol = a.t.qtol[idx];
        sThr[idx] = tol < 0.f ? 1e30f : 0.5f - tol;
        sZz[idx] = a.t.zz[idx];
    }
    if (lane == 0) {
        big[0] = 0;
        *rq_count = 0;
        ff_mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tables_const) jb_pdl_wait();

    const int P = jb_ctrl_parity(a);
    unsigned* const ticket = jb_ctrl_ticket(a, P);
    const bool refine_on = !(g.flags & JB_FLAG_NO_REFINE);
    const bool use_tma = ka.use_tma != 0;
    const int hb = g.hb;
    const int rg = lane >> 2, kg = lane & 3, kpar = kg >> 1, khalf = kg & 1;

    // first chunk of every warp dealt statically, CTA-major, so that a single frame spreads over all SMs; the rest
    // from the ticket counter (as in jb_forward_fast.cu)
    const unsigned static_chunks = gridDim.x * (blockDim.x >> 5);
    if (blockIdx.x == 0 && threadIdx.x == 0) a.ctrl[1] = (unsigned)P;
    auto claim = [&]() -> unsigned {
        unsigned c = 0;
        if (lane == 0) c = static_chunks + atomicAdd(ticket, 1u);
        return __shfl_sync(0xffffffffu, c, 0);
    };
    // a block whose whole tile lies inside the image can come by TMA
    auto interior = [&](int by, int bx) -> bool { return (by + 1) * side <= g.H && (bx + 1) * side <= g.W; };
    bool dirty = true;                                   // generic-proxy writes to the tile since the last TMA load
    auto issue = [&](int plane, int by, int bx) {
        if (lane == 0) {
            if (dirty) ff_fence_proxy_async();
            ff_mbar_expect_tx(bar, (uint32_t)(side * FL_TILE_ROW));
            ff_tma_load_3d(tile, &tmap, (bx * side) & ~15, by * side, plane, bar);      // (16-byte aligned start)
        }
    };
    uint32_t phase = 0;

    unsigned chunk = blockIdx.x + gridDim.x * (unsigned)warp;
    // (plane, blk0, nvalid) of the chunk, and the state of the tile slot: has the load of the chunk's first block
    // been issued already (by the previous chunk's last block)?
    bool first_in_flight = false;
    if (chunk < a.n_chunks && use_tma) {
        const int plane = (int)(chunk / (unsigned)g.cpp);
        const int blk0 = (int)(chunk - (unsigned)plane * (unsigned)g.cpp) * g.chunk;
        const int by = blk0 / hb, bx = blk0 - by * hb;
        if (interior(by, bx)) { issue(plane, by, bx); first_in_flight = true; }
        dirty = false;
    }
    while (chunk < a.n_chunks) {
        const int plane = (int)(chunk / (unsigned)g.cpp);
        const int blk0 = (int)(chunk - (unsigned)plane * (unsigned)g.cpp) * g.chunk;
        const int nvalid = jb_min(g.chunk, g.nblocks - blk0);
        const uint8_t* src = a.planes + (size_t)plane * a.plane_stride;
        uint8_t* slot = a.tmp + (size_t)chunk * a.chunk_cap;           // (the large path always uses the big slots)
        unsigned chunk_bytes = 0;
        unsigned next_chunk = 0;
        bool in_flight = first_in_flight;
        first_in_flight = false;

        int by = blk0 / hb, bx = blk0 - by * hb;
        for (int gi = 0; gi < nvalid; ++gi, ++bx) {
            const int blk = blk0 + gi;
            if (bx >= hb) { bx = 0; ++by; }
            // the ticket of the chunk after this one is drawn a block early: the round trip of the atomic hides behind
            // this block's box sums
            unsigned early_ticket = 0;
            if (gi + 1 == nvalid && lane == 0) early_ticket = static_chunks + atomicAdd(ticket, 1u);
            // ---- the tile of this block; xoff: where the block's first column sits in the tile rows ----
            int xoff = 0;
            if (in_flight) {
                while (!ff_mbar_try_wait(bar, phase)) { }
                phase ^= 1u;
                xoff = (bx * side) & 15;
            } else {
                // edge block, or no TMA: the reference's two-level edge replication, byte by byte
                for (int idx = lane; idx < side * side; idx += 32) {
                    const int r = idx / side, c = idx - r * side;
                    const int si = jb_min(by * D + r / bs, g.H1 - 1), sj = jb_min(bx * D + c / bs, g.W1 - 1);
                    const int y = jb_min(si * bs + r % bs, g.H - 1), x = jb_min(sj * bs + c % bs, g.W - 1);
                    tile[r * FL_TILE_ROW + c] = src[(size_t)y * a.row_pitch + x];
                }
                dirty = true;
                __syncwarp();
            }
            // ---- box sums -> Xe, Xo.  Lane j < D sums sample (i, j): bs rows x bs bytes from byte bs j, under byte masks ----
            {
                const int j = lane < D ? lane : D - 1;
                const int x0 = xoff + j * bs;
                const int w0 = x0 >> 2, o0 = x0 & 3;
                // masks of the up to three words a bs-byte run (bs <= 8) touches
                const int nb0 = jb_min(4 - o0, bs);                               // bytes of the run in word 0
                const uint32_t m0 = (nb0 >= 4 ? 0x01010101u : (0x01010101u >> (8 * (4 - nb0)))) << (8 * o0);
                const int nb1 = jb_min(4, bs - nb0);
                const uint32_t m1 = nb1 <= 0 ? 0u : (nb1 >= 4 ? 0x01010101u : (0x01010101u >> (8 * (4 - nb1))));
                const int nb2 = bs - nb0 - (nb1 > 0 ? nb1 : 0);
                const uint32_t m2 = nb2 <= 0 ? 0u : (0x01010101u >> (8 * (4 - nb2)));
                const uint32_t* t32 = (const uint32_t*)tile + w0;
                const int partner = D - 1 - lane;                                  // (lanes < D)
                constexpr int RW = FL_TILE_ROW / 4;                                // words per tile row
                if constexpr (BS == 5 && D == 24) {
                    // config 3 (side 120, so xoff is 0 or 8): lane i < D sums sample row i -- tile rows 5 i .. 5 i + 4, each
                    // eight 128-bit loads and 24 x 2 dot products under compile-time byte masks into 24 accumulators --
                    // and forms the even / odd sums of its row without leaving its registers.  Rows are 144 bytes apart:
                    // the eight lanes of a load phase fall on distinct bank groups.
                    if (lane < D) {
                        uint32_t acc[D];
                        #pragma unroll
                        for (int jj = 0; jj < D; ++jj) acc[jj] = 0x4B000000u;     // bit pattern of 2^23: exact int -> float below
                        const uint4* rows = (const uint4*)(tile + (size_t)(5 * lane) * FL_TILE_ROW);
                        auto sum_rows = [&](auto xw_tag) {
                            constexpr int XW = decltype(xw_tag)::value;          // xoff in words
                            #pragma unroll
                            for (int k = 0; k < 5; ++k) {
                                uint32_t w[32];
                                #pragma unroll
                                for (int c = 0; c < 8; ++c) {
                                    const uint4 v = rows[k * (FL_TILE_ROW / 16) + c];
                                    w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
                                }
                                #pragma unroll
                                for (int jj = 0; jj < D; ++jj) {
                                    const int b0 = 4 * XW + 5 * jj, wi = b0 >> 2, o = b0 & 3;
                                    acc[jj] = __dp4a(w[wi], 0x01010101u << (8 * o), acc[jj]);               // bytes o..3
                                    acc[jj] = __dp4a(w[wi + 1], 0x01010101u >> (8 * (3 - o)), acc[jj]);     // bytes 0..o
                                }
                            }
                        };
                        if (xoff == 0) sum_rows(std::integral_constant<int, 0>()); else sum_rows(std::integral_constant<int, 2>());
                        float e[H], od[H];
                        #pragma unroll
                        for (int jj = 0; jj < H; ++jj) {
                            const float f = __uint_as_float(acc[jj]) - 8388608.0f, p = __uint_as_float(acc[D - 1 - jj]) - 8388608.0f;
                            e[jj] = f + p; od[jj] = f - p;
                        }
                        float4* de = (float4*)(sXe + lane * L.xs);
                        float4* dd = (float4*)(sXo + lane * L.xs);
                        #pragma unroll
                        for (int q4 = 0; q4 < H / 4; ++q4) {
                            de[q4] = make_float4(e[4 * q4], e[4 * q4 + 1], e[4 * q4 + 2], e[4 * q4 + 3]);
                            dd[q4] = make_float4(od[4 * q4], od[4 * q4 + 1], od[4 * q4 + 2], od[4 * q4 + 3]);
                        }
                    }
                } else if (BS > 0 && BS <= 5) {
                    // compile-time block_size whose runs touch two words at most: everything unrolled, four sample rows
                    // (8 BS independent loads, 4 accumulator chains) in flight at a time
                    #pragma unroll
                    for (int i0 = 0; i0 < D; i0 += 4) {
                        uint32_t acc[4];
                        #pragma unroll
                        for (int ii = 0; ii < 4; ++ii) {
                            const uint32_t* row = t32 + (i0 + ii) * BS * RW;
                            uint32_t sa = 0x4B000000u, sb = 0u;    // 0x4B000000 + s is the bit pattern of 2^23 + s: exact int -> float
                            #pragma unroll
                            for (int k = 0; k < BS; ++k) {
                                sa = __dp4a(row[k * RW], m0, sa);
                                sb = __dp4a(row[k * RW + 1], m1, sb);
                            }
                            acc[ii] = sa + sb;
                        }
                        #pragma unroll
                        for (int ii = 0; ii < 4; ++ii) {
                            const float f = __uint_as_float(acc[ii]) - 8388608.0f;
                            const float p = __shfl_sync(0xffffffffu, f, partner & 31);
                            if (lane < H) { sXe[(i0 + ii) * L.xs + lane] = f + p; sXo[(i0 + ii) * L.xs + lane] = f - p; }
                        }
                    }
                } else {
                    #pragma unroll 2
                    for (int i = 0; i < D; ++i) {
                        const uint32_t* row = t32 + (size_t)i * bs * RW;
                        uint32_t sacc = 0x4B000000u;
                        for (int k = 0; k < bs; ++k) {
                            sacc = __dp4a(row[0], m0, sacc);
                            sacc = __dp4a(row[1], m1, sacc);
                            if (m2) sacc = __dp4a(row[2], m2, sacc);
                            row += RW;
                        }
                        const float f = __uint_as_float(sacc) - 8388608.0f;
                        const float p = __shfl_sync(0xffffffffu, f, partner & 31);
                        if (lane < H) { sXe[i * L.xs + lane] = f + p; sXo[i * L.xs + lane] = f - p; }
                    }
                }
            }
            __syncwarp();
            // ---- the tile is consumed: the next block's tile may come in (this chunk's, or the next chunk's first) ----
            in_flight = false;
            {
                int nplane = plane, nblk = blk + 1;
                bool have = gi + 1 < nvalid;
                if (!have) {
                    next_chunk = __shfl_sync(0xffffffffu, early_ticket, 0);
                    if (next_chunk < a.n_chunks) {
                        nplane = (int)(next_chunk / (unsigned)g.cpp);
                        nblk = (int)(next_chunk - (unsigned)nplane * (unsigned)g.cpp) * g.chunk;
                        have = true;
                    }
                }
                if (have && use_tma) {
                    const int nby = nblk / hb, nbx = nblk - nby * hb;
                    if (interior(nby, nbx)) {
                        issue(nplane, nby, nbx);
                        dirty = false;
                        if (gi + 1 < nvalid) in_flight = true; else first_in_flight = true;
                    }
                }
            }
            // ---- row pass: T[i][k] = sum_{j < H} (k even ? Xe : Xo)[i][j] C[k][j]; lane tile RT rows x KT frequencies ----
            float t[RT][KT];
            {
                const float* xbase = (kpar ? sXo : sXe) + (rg * RT) * L.xs;
                const float* cbase = sCH + (kpar * H + khalf * KT) * L.xs;
                #pragma unroll
                for (int r = 0; r < RT; ++r)
                    #pragma unroll
                    for (int q = 0; q < KT; ++q) t[r][q] = 0.f;
                #pragma unroll
                for (int j = 0; j < H; j += 4) {
                    float4 x4[RT], c4[KT];
                    #pragma unroll
                    for (int r = 0; r < RT; ++r) x4[r] = *(const float4*)(xbase + r * L.xs + j);
                    #pragma unroll
                    for (int q = 0; q < KT; ++q) c4[q] = *(const float4*)(cbase + q * L.xs + j);
                    #pragma unroll
                    for (int r = 0; r < RT; ++r)
                        #pragma unroll
                        for (int q = 0; q < KT; ++q) {
                            t[r][q] = fmaf(x4[r].x, c4[q].x, t[r][q]); t[r][q] = fmaf(x4[r].y, c4[q].y, t[r][q]);
                            t[r][q] = fmaf(x4[r].z, c4[q].z, t[r][q]); t[r][q] = fmaf(x4[r].w, c4[q].w, t[r][q]);
                        }
                }
            }
            // ---- even / odd sums down the columns: rows i and D-1-i sit in lanes rg and 7 - rg (same kg) ----
            {
                float other[RT][KT];
                #pragma unroll
                for (int r = 0; r < RT; ++r)
                    #pragma unroll
                    for (int q = 0; q < KT; ++q) other[RT - 1 - r][q] = __shfl_xor_sync(0xffffffffu, t[r][q], 28);
                // lanes rg < 4 hold rows i < H and store Te[i] = T[i] + T[D-1-i]; lanes rg >= 4 hold rows D-1-i' and
                // store To[i'] = T[i'] - T[D-1-i'] = other - mine
                const bool low = rg < 4;
                float* dstT = low ? sTe : sTo;
                #pragma unroll
                for (int r = 0; r < RT; ++r) {
                    const int i = low ? rg * RT + r : D - 1 - (rg * RT + r);
                    #pragma unroll
                    for (int q = 0; q < KT; ++q) {
                        const int k = 2 * (khalf * KT + q) + kpar;
                        dstT[i * L.ts + k] = low ? t[r][q] + other[r][q] : other[r][q] - t[r][q];
                    }
                }
            }
            __syncwarp();
            // ---- column pass: Y[u][k] = sum_{i < H} (u even ? Te : To)[i][k] C[u][i]; lane tile RT frequencies u (one
            //      parity) x KT consecutive k; quantise, tie check, zigzag ----
            {
                const int upar = rg >> 2, ug = rg & 3;                  // u = 2 (ug RT + r) + upar
                const float* tb = (upar ? sTo : sTe) + kg * KT;
                const float* cb = sCH + (upar * H + ug * RT) * L.xs;
                float y[RT][KT];
                #pragma unroll
                for (int r = 0; r < RT; ++r)
                    #pragma unroll
                    for (int q = 0; q < KT; ++q) y[r][q] = 0.f;
                #pragma unroll
                for (int i = 0; i < H; i += 4) {
                    float4 c4[RT];
                    #pragma unroll
                    for (int r = 0; r < RT; ++r) c4[r] = *(const float4*)(cb + r * L.xs + i);
                    #pragma unroll
                    for (int ii = 0; ii < 4; ++ii) {
                        float tv[KT];
                        #pragma unroll
                        for (int q = 0; q < KT; q += 2) {
                            const float2 v2 = *(const float2*)(tb + (i + ii) * L.ts + q);
                            tv[q] = v2.x; tv[q + 1] = v2.y;
                        }
                        #pragma unroll
                        for (int r = 0; r < RT; ++r) {
                            const float cv = ii == 0 ? c4[r].x : ii == 1 ? c4[r].y : ii == 2 ? c4[r].z : c4[r].w;
                            #pragma unroll
                            for (int q = 0; q < KT; ++q) y[r][q] = fmaf(cv, tv[q], y[r][q]);
                        }
                    }
                }
                __syncwarp();                    // (the quantised row shares its memory with Te / To: every lane has read them)
                // quantise all RT x KT outputs straight through (no branches); the rare cases -- a value within the fp32
                // error bound of a rounding tie, an amplitude beyond the 15-bit size field -- are collected in bit masks
                // and dealt with behind the loop
                unsigned nearmask = 0, bigmask = 0;
                int16_t* const cout = MODE == 1 ? a.coeffs_out + ((size_t)plane * g.nblocks + blk) * n : crow;
                float vmax = 0.f;
                #pragma unroll
                for (int r = 0; r < RT; ++r) {
                    const int u = 2 * (ug * RT + r) + upar;
                    const int idx0 = u * D + kg * KT;                   // (even: 8-byte aligned table rows)
                    #pragma unroll
                    for (int q = 0; q < KT; q += 2) {
                        const float2 qm2 = *(const float2*)(sQm + idx0 + q), th2 = *(const float2*)(sThr + idx0 + q);
                        const uint32_t zz2 = *(const uint32_t*)(sZz + idx0 + q);
                        #pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const float val = y[r][q + h] * (h ? qm2.y : qm2.x);
                            vmax = fmaxf(vmax, fabsf(val));
                            const float tt = val + 12582912.0f;         // 1.5 * 2^23: rounds half-even to an integer
                            const int qi = __float_as_int(tt) - 0x4B400000;
                            const float dd = fabsf(val - (tt - 12582912.0f));
                            if (fmaf(fabsf(val), 2.4e-7f, dd) > (h ? th2.y : th2.x)) nearmask |= 1u << (r * KT + q + h);
                            cout[h ? zz2 >> 16 : zz2 & 0xFFFFu] = (int16_t)qi;
                            y[r][q + h] = __int_as_float(qi);           // (kept for the rare case below)
                        }
                    }
                }
                if (vmax > (float)JB_MAX_AMP - 0.75f) {
                    // some amplitude may not fit the 15-bit size field: store it saturated, and see below
                    #pragma unroll
                    for (int r = 0; r < RT; ++r)
                        #pragma unroll
                        for (int q = 0; q < KT; ++q) {
                            const int qi = __float_as_int(y[r][q]);
                            if (qi > JB_MAX_AMP || qi < -JB_MAX_AMP) {
                                bigmask |= 1u << (r * KT + q);
                                cout[sZz[(2 * (ug * RT + r) + upar) * D + kg * KT + q]] = (int16_t)max(-32767, min(32767, qi));
                            }
                        }
                }
                if (!refine_on) nearmask = 0;
                if (nearmask | bigmask) {
                    for (unsigned mm = nearmask | bigmask; mm; mm &= mm - 1u) {
                        const int t = __ffs((int)mm) - 1, r = t / KT, q = t - r * KT;
                        const int idx = (2 * (ug * RT + r) + upar) * D + kg * KT + q;
                        if (nearmask >> t & 1u) {
                            // queued for the warp's float64 re-evaluation below (a lane on its own would take ~30 times as
                            // long while 31 lanes wait for it)
                            const int slot = atomicAdd(rq_count, 1);
                            if (slot < FL_RQ_CAP) { rq[slot] = (uint16_t)idx; continue; }
                            const int qi = (int)rint(fl_refine<D>(sXe, idx, (bs << 8) | g.qmode, a.t.fA64, a.t.qrecip[idx]));
                            cout[sZz[idx]] = (int16_t)max(-32767, min(32767, qi));
                            if (MODE != 1 && (qi > JB_MAX_AMP || qi < -JB_MAX_AMP)) {
                                const int kk = atomicAdd(big, 1);
                                if (kk < FL_BIG_CAP) { big[1 + 2 * kk] = sZz[idx]; big[2 + 2 * kk] = qi; }
                            }
                        } else if (MODE != 1) {
                            // BadRleCodeError in the reference (util.py:170-171): remember the true amplitude for the report
                            // the run-length stage makes (the stored value is saturated)
                            int qi = 0;
                            #pragma unroll
                            for (int rr = 0; rr < RT; ++rr)
                                #pragma unroll
                                for (int qq = 0; qq < KT; ++qq) if (rr * KT + qq == t) qi = __float_as_int(y[rr][qq]);
                            const int kk = atomicAdd(big, 1);
                            if (kk < FL_BIG_CAP) { big[1 + 2 * kk] = sZz[idx]; big[2 + 2 * kk] = qi; }
                        }
                    }
                }
            }
            __syncwarp();
            // ---- float64 re-evaluation of the queued coefficients, the whole warp on one coefficient at a time, in the
            //      reference's order (jb_refine.cuh): lane i sums row i against C[v] (ascending j, separate multiply and
            //      add), then the 24 row results are combined against C[u] in ascending i ----
            {
                const int nreq = jb_min(*rq_count, FL_RQ_CAP);
                if (nreq > 0) {
                    const JbBoxMean mean(bs);
                    const double* A64 = a.t.fA64;
                    const int li = lane < D ? lane : D - 1;
                    for (int t = 0; t < nreq; ++t) {
                        const int idx = rq[t];
                        const int u = idx / D, v = idx - u * D;
                        double m = 0.0;
                        for (int j = 0; j < D; ++j) {
                            const int jj = j < H ? j : D - 1 - j;
                            const float e = sXe[li * L.xs + jj], o = sXo[li * L.xs + jj];
                            const double x = mean((double)((j < H ? e + o : e - o) * 0.5f));
                            const double term = __dmul_rn(x, A64[v * D + j]);
                            m = j == 0 ? term : __dadd_rn(m, term);
                        }
                        double y64 = 0.0;
                        for (int i = 0; i < D; ++i) {
                            const double term = __dmul_rn(__shfl_sync(0xffffffffu, m, i), A64[u * D + i]);
                            y64 = i == 0 ? term : __dadd_rn(y64, term);
                        }
                        const double recip = a.t.qrecip[idx];
                        const double f64 = g.qmode == JB_Q_QTABLE ? __dmul_rn(y64, recip) : g.qmode == JB_Q_DIVIDE ? __ddiv_rn(y64, recip) : y64;
                        if (lane == 0) {
                            int qi = (int)rint(f64);
                            const int zp = sZz[idx];
                            if (MODE == 1) {
                                a.coeffs_out[((size_t)plane * g.nblocks + blk) * n + zp] = (int16_t)max(-32767, min(32767, qi));
                            } else {
                                if (qi > JB_MAX_AMP || qi < -JB_MAX_AMP) {
                                    const int kk = atomicAdd(big, 1);
                                    if (kk < FL_BIG_CAP) { big[1 + 2 * kk] = zp; big[2 + 2 * kk] = qi; }
                                    qi = qi > 0 ? 32767 : -32767;
                                }
                                crow[zp] = (int16_t)qi;
                            }
                        }
                    }
                }
                if (lane == 0) *rq_count = 0;
                __syncwarp();
            }
            if (MODE == 1) continue;
            // ---- non-zero bitmap (zigzag order): lane w keeps word w ----
            uint32_t mybits = 0;
            #pragma unroll
            for (int w = 0; w < NW; ++w) {
                const uint32_t m = __ballot_sync(0xffffffffu, crow[w * 32 + lane] != 0);
                if (lane == w) mybits = m;
            }
            // ---- run-length codes by the whole warp, 32 zigzag positions at a time (empty words cost nothing): lane l
            //      of non-empty word w looks at position 32 w + l; the previous non-zero is the nearest set bit below it
            //      (or the last one of the words before), a warp scan of the code lengths gives the bit offsets, and the
            //      codes are OR-ed into the block's bytes in shared memory ----
            for (int i = lane; i < FL_BLOCK_BUF_WORDS + 2; i += 32) obuf[i] = 0u;
            __syncwarp();
            unsigned bitpos = 0;                       // bits emitted so far (warp-uniform)
            int prev = -1;                             // last non-zero position of the words before (warp-uniform)
            bool any_bad = false, overflow = false;
            for (uint32_t wleft = __ballot_sync(0xffffffffu, mybits != 0); wleft; wleft &= wleft - 1u) {
                const int w = __ffs((int)wleft) - 1;
                const uint32_t word = __shfl_sync(0xffffffffu, mybits, w);
                const int p = w * 32 + lane;
                const bool nz = (word >> lane) & 1u;
                const uint32_t below = word & ((1u << lane) - 1u);
                const int pv = below ? w * 32 + 31 - __clz((int)below) : prev;
                const int amp = nz ? (int)crow[p] : 0;
                const uint32_t mag = (uint32_t)(amp < 0 ? -amp : amp);
                int run = p - pv - 1;
                const bool bad = nz && mag > (uint32_t)JB_MAX_AMP;        // BadRleCodeError: nothing is emitted (as in the reference)
                const int size = 33 - __clz((int)(mag | 1u));             // bit length + 1
                const unsigned nzrl = (unsigned)(run / JB_MAX_RUN);
                const unsigned len = (nz && !bad) ? 8u * nzrl + 8u + (unsigned)size : 0u;
                unsigned total;
                unsigned off = bitpos + jb_warp_excl_scan(len, lane, &total);
                if (bitpos + total + 8u > FL_BLOCK_BUF_WORDS * 32u) { overflow = true; break; }      // (warp-uniform)
                if (bad) {
                    long long true_amp = amp;
                    const int nb = jb_min(big[0], FL_BIG_CAP);
                    for (int k = 0; k < nb; ++k)
                        if (big[1 + 2 * k] == p) true_amp = big[2 + 2 * k];
                    jb_report_bad_code(jb_ctrl_status(a, P), (unsigned long long)plane * g.nblocks + blk, p, run % JB_MAX_RUN, true_amp);
                }
                if (len) {
                    for (unsigned z = 0; z < nzrl; ++z) {                       // (15, 0): the byte 0xF0
                        atomicOr(obuf + (off >> 5), 0xF0000000u >> (off & 31u));
                        if ((off & 31u) > 24u) atomicOr(obuf + (off >> 5) + 1, 0xF0000000u << (32u - (off & 31u)));
                        off += 8u;
                    }
                    run -= (int)nzrl * JB_MAX_RUN;
                    const uint32_t code = ((((uint32_t)run << 4) | (uint32_t)size) << size) | ((amp > 0 ? 1u : 0u) << (size - 1)) | mag;
                    const unsigned cl = 8u + (unsigned)size;
                    const unsigned long long v64 = (unsigned long long)code << (64u - cl - (off & 31u));
                    atomicOr(obuf + (off >> 5), (uint32_t)(v64 >> 32));
                    if ((uint32_t)v64) atomicOr(obuf + (off >> 5) + 1, (uint32_t)v64);
                }
                any_bad = any_bad || __any_sync(0xffffffffu, bad);
                bitpos += total;
                prev = w * 32 + 31 - __clz((int)word);
            }
            uint8_t* dst = slot + chunk_bytes;
            unsigned blen;
            if (!overflow) {
                blen = (bitpos + 8u + 7u) >> 3;                                 // + EOB, padded to the byte
                __syncwarp();
                // bytes of the block -> the chunk's slot (big-endian words = stream order)
                for (unsigned i = lane; i < blen; i += 32) dst[i] = (uint8_t)(obuf[i >> 2] >> (24u - 8u * (i & 3u)));
            } else {
                // a block longer than the shared-memory buffer: lane 0 packs it serially, straight to the slot (amplitudes
                // that do not fit are skipped there as well; they are reported by the code above only up to the overflow,
                // so report the rest here)
                __shared__ uint32_t s_mask[FL_WARPS][32];
                if (lane < NW) s_mask[warp][lane] = mybits;
                __syncwarp();
                unsigned bl = 0;
                if (lane == 0) {
                    bl = fl_pack_serial(crow, s_mask[warp], NW, dst);
                    int pr = -1;
                    for (int wi = 0; wi < NW; ++wi)
                        for (uint32_t m = s_mask[warp][wi]; m; m &= m - 1u) {
                            const int p = wi * 32 + __ffs((int)m) - 1;
                            const int amp = crow[p];
                            if (amp > JB_MAX_AMP || amp < -JB_MAX_AMP) {
                                long long true_amp = amp;
                                const int nb = jb_min(big[0], FL_BIG_CAP);
                                for (int k = 0; k < nb; ++k)
                                    if (big[1 + 2 * k] == p) true_amp = big[2 + 2 * k];
                                jb_report_bad_code(jb_ctrl_status(a, P), (unsigned long long)plane * g.nblocks + blk, p,
                                                   (p - pr - 1) % JB_MAX_RUN, true_amp);
                                any_bad = true;
                            }
                            pr = p;
                        }
                }
                blen = __shfl_sync(0xffffffffu, bl, 0);
                any_bad = __any_sync(0xffffffffu, any_bad);
            }
            chunk_bytes += blen;
            if (any_bad && lane == 0) big[0] = 0;
            __syncwarp();
        }
        if (MODE == 0 && lane == 0) jb_record_chunk_len(a, P, chunk, chunk_bytes);
        if (nvalid < 1) next_chunk = claim();               // (cannot happen: every chunk holds a block)
        chunk = next_chunk;
    }
}

// ---- host side ----------------------------------------------------------------------------------
template <int D, int MODE, int BS>
static cudaError_t fl_launch_t(const CUtensorMap& map, const FlKernelArgs& ka, cudaStream_t s) {
    const size_t smem = fl_layout(D, D * ka.a.g.bs).total;
    cudaError_t e = cudaFuncSetAttribute(jb_fwd_large_kernel<D, MODE, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned want = ka.a.n_chunks;
    const unsigned grid = want < (unsigned)sms ? want : (unsigned)sms;
    if (grid == 0) return cudaSuccess;
    return jb_launch_ex(jb_fwd_large_kernel<D, MODE, BS>, dim3(grid), dim3(fl_layout(D, D * ka.a.g.bs).warps * 32), smem, s,
                        (ka.a.g.flags & JB_FLAG_PDL) != 0, map, ka);
}

template <int MODE>
static cudaError_t fl_launch_d(const CUtensorMap& map, const FlKernelArgs& ka, cudaStream_t s) {
    switch (ka.a.g.d) {
    case 16: return fl_launch_t<16, MODE, 0>(map, ka, s);
    case 24: return ka.a.g.bs == 5 ? fl_launch_t<24, MODE, 5>(map, ka, s) : fl_launch_t<24, MODE, 0>(map, ka, s);
    default: return fl_launch_t<32, MODE, 0>(map, ka, s);
    }
}

cudaError_t jb_launch_fwd_large(const JbFwdArgs& a_in, int mode, cudaStream_t s) {
    FlKernelArgs ka;
    ka.a = a_in;
    ka.a.slots_always_big = 1;
    const JbGeom& g = ka.a.g;
    const int side = g.d * g.bs;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    const bool aligned = ((uintptr_t)ka.a.planes & 15) == 0 && (ka.a.row_pitch & 15) == 0 &&
                         (ka.a.n_planes == 1 || (ka.a.plane_stride & 15) == 0);
    ka.use_tma = 0;
    if (!(g.flags & JB_FLAG_NO_TMA) && aligned)
        ka.use_tma = jb_make_plane_tensor_map(&map, ka.a.planes, g.W, g.H, ka.a.n_planes, ka.a.row_pitch, ka.a.plane_stride,
                                              FL_TILE_ROW, side) ? 1 : 0;
    if (mode == 0) {
        cudaError_t e = fl_launch_d<0>(map, ka, s);
        if (e != cudaSuccess) return e;
        return jb_launch_gather(ka.a, s);
    }
    return fl_launch_d<1>(map, ka, s);
}
