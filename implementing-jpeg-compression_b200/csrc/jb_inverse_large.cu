// jb_inverse_large.cu -- decompress direction for large blocks: DCT, dct_size 16 / 24 / 32, tile rows of at most
// 128 bytes (BASELINE.json config 3: --block_size 5 --dct_size 24 --quantization divide --qdivisor 1000).
//
// The mirror of jb_forward_large.cu.  A warp of a persistent grid claims a chunk of g.chunk = 8 (small calls: 2) consecutive
// blocks of one plane (block offsets come from jb_framing.cu) and takes it through every stage on its own:
//   * eight lanes decode the eight blocks (rle_byte_stream.py:74-88, run_length_encoding.py:31-41) into natural
//     (un-zigzagged, zigzag_order.py:101-119) int16 rows in shared memory;
//   * per block: dequantise (quantizers.py restore) and X = B.Y.B^T (transforms.py:60-69) as two half-length
//     contractions: B[d-1-m][k] = (-1)^k B[m][k], so the even and the odd frequencies are summed separately for the
//     first d/2 samples only, and sample m / d-1-m are their sum / difference -- d/2 instead of d multiply-adds per
//     output.  Lane (rg, kg) owns a (d/8) x (d/4) register tile; operands come from shared memory as 128-bit rows;
//     the two halves of a sum sit in partner lanes (__shfl_xor 2 along a row, 16 down a column);
//   * np.round, clamp (basis_change.py:43, normalization.py:10-14) -> 8-bit samples of the chunk in shared memory,
//     laid out sample row by sample row across the chunk's blocks;
//   * the chunk goes out pixel row by pixel row (util.inflate, subsampling.py:13-14): a run of 8 d bs contiguous bytes
//     per row, 128-bit stores, each vector built from <= 8 samples with four byte permutes (selectors fixed per lane)
//     and written to the bs rows it covers.  Chunks that wrap a block row, touch the cropped edges
//     (dct_padding.py:11-21, padding.py:14-16) or are not 16-byte aligned use byte stores.
#include <string.h>

#include "jb_common.cuh"
#include "jb_inverse.cuh"

#define IL_MAX_WARPS 10

struct IlLayout {
    int h;
    size_t bh, dq, izz;                                   // CTA-wide tables
    size_t warp0, coef, y, p, samp, warp_bytes, total;
    int warps;
};

__host__ __device__ inline IlLayout il_layout(int d) {
    IlLayout L;
    const int n = d * d;
    L.h = d / 2;
    size_t o = 0;
    L.bh = o;  o += (size_t)n * 4 / 2;                    // BH[parity][c < h][m < h]
    L.dq = o;  o += (size_t)n * 4;
    L.izz = o; o += jb_align_up((size_t)n * 2, 16);
    L.warp0 = jb_align_up(o, 128);
    size_t w = 0;
    L.coef = w; w += (size_t)JB_CHUNK_LARGE * n * 2;
    L.y = w;    w += (size_t)n * 4;                       // Ye[u][m], then Yo[u][m]
    L.p = w;    w += (size_t)n * 4;                       // Pev[mu][c], then Pod[mu][c]
    L.samp = w; w += (size_t)JB_CHUNK_LARGE * n + 16;     // [sample row][block of the chunk][sample column]
    L.warp_bytes = jb_align_up(w, 128);
    int warps = (int)((220 * 1024 - L.warp0) / L.warp_bytes);
    L.warps = warps > IL_MAX_WARPS ? IL_MAX_WARPS : warps;
    L.total = L.warp0 + (size_t)L.warps * L.warp_bytes;
    return L;
}

bool jb_inv_large_eligible(const JbGeom& g) {
    if (g.transform != JB_TRANSFORM_DCT) return false;
    if (g.d != 16 && g.d != 24 && g.d != 32) return false;
    return g.d * g.bs <= 128 && il_layout(g.d).warps >= 4;
}

// word-based block decoder over global memory, any n (same logic as fi_decode_block of jb_inverse_fast.cu); also
// reports the largest row / column (natural order) that holds a coefficient: umax, vmax (-1: the block is empty)
template <int D>
__device__ __forceinline__ int il_decode_block(const uint8_t* stream, uint32_t start, uint32_t len, int n,
                                               int16_t* row, const uint16_t* izz, int& umax, int& vmax) {
    umax = vmax = -1;
    const uintptr_t a0 = (uintptr_t)(stream + start);
    const uint32_t mis = (uint32_t)(a0 & 3);
    const uint32_t* words = (const uint32_t*)(a0 - mis);
    const uint32_t nwords = (len - start + mis + 3u) >> 2;
    const uint32_t bitlimit = (len - start + mis) * 8u;
    uint32_t widx = 0, used = mis * 8u;
    // the first IL_PRE words in one round trip (a heavily quantised block is 15-30 bytes: config 3), the rest on demand
    constexpr uint32_t IL_PRE = 8;
    uint32_t pw[IL_PRE];
    #pragma unroll
    for (uint32_t q = 0; q < IL_PRE; ++q) pw[q] = q < nwords ? __ldg(words + q) : 0u;
    auto ld = [&](uint32_t i) -> uint32_t {
        if (i < IL_PRE) return jb_bswap32(pw[i]);
        return i < nwords ? jb_bswap32(__ldg(words + i)) : 0u;
    };
    uint64_t buf = ld(widx++);
    int nb = 32 - (int)used;
    int count = 0;
    for (;;) {
        if (nb < 23) { buf = (buf << 32) | ld(widx++); nb += 32; }
        const uint32_t head = (uint32_t)(buf >> (nb - 8)) & 0xFFu;
        const uint32_t run = head >> 4, size = head & 15u;
        if (size == 0u) {
            nb -= 8; used += 8;
            if (run == 0u) return used > bitlimit ? 1 : 0;
            if (run != (uint32_t)JB_MAX_RUN) return 1;
            count += JB_MAX_RUN;
            if (count > n || used > bitlimit) return 1;
            continue;
        }
        if (size == 1u) return 1;
        count += (int)run;
        if (count >= n) return 1;
        const uint32_t raw = (uint32_t)(buf >> (nb - 8 - (int)size)) & ((1u << size) - 1u);
        nb -= 8 + (int)size; used += 8u + size;
        if (used > bitlimit) return 1;
        const int mag = (int)(raw & ((1u << (size - 1)) - 1u));
        const int nat = izz[count];
        row[nat] = (int16_t)((raw >> (size - 1)) ? mag : -mag);
        umax = max(umax, nat / D);
        vmax = max(vmax, nat % D);
        ++count;
    }
}

// BS: block_size as a compile-time constant (5: config 3), so that the row stores unroll; 0: any block_size at run time.
template <int D, int MODE, int BS>
__global__ void __launch_bounds__(IL_MAX_WARPS * 32, 1)
jb_inv_large_kernel(const JbInvArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const JbGeom& g = a.g;
    constexpr int n = D * D, H = D / 2, RT = D / 8, KT = D / 4, ROWB = JB_CHUNK_LARGE * D;     // bytes per sample row of a chunk
    const int bs = BS ? BS : g.bs, side = D * bs;
    const IlLayout L = il_layout(D);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NWARPS = blockDim.x >> 5;
    float* sBH = (float*)(smem + L.bh);
    float* sDq = (float*)(smem + L.dq);
    uint16_t* sIzz = (uint16_t*)(smem + L.izz);
    unsigned char* wbase = smem + L.warp0 + (size_t)warp * L.warp_bytes;
    int16_t* coef = (int16_t*)(wbase + L.coef);
    float* sY = (float*)(wbase + L.y);
    float* sP = (float*)(wbase + L.p);
    uint8_t* samp = wbase + L.samp;

    jb_pdl_trigger();
    const bool tables_const = (g.flags & JB_FLAG_REUSE_TABLES) != 0;
    if (!tables_const) jb_pdl_wait();
    // BH[par][c][m] = B[c][2 m + par] for the first H samples c; B[c][k] = iA[c * D + k] (sample c, frequency k)
    for (int idx = threadIdx.x; idx < 2 * H * H; idx += blockDim.x) {
        const int par = idx / (H * H), rem = idx - par * H * H, c = rem / H, m = rem - c * H;
        sBH[idx] = a.t.iA[c * D + 2 * m + par];
    }
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        sDq[idx] = a.t.dqmult[idx];
        sIzz[idx] = a.t.izz[idx];
    }
    __syncthreads();
    if (tables_const) jb_pdl_wait();

    const int rg = lane >> 2, kg = lane & 3;
    const int hb = g.hb;
    uint8_t* const plane0 = a.planes_out;
    const bool aligned = (((uintptr_t)plane0 & 15) == 0) && ((a.row_pitch & 15) == 0) &&
                         (a.n_planes == 1 || (a.plane_stride & 15) == 0);
    // byte-permute selectors of this lane's (up to two) 16-byte vectors of a pixel row: output byte t of vector v
    // shows sample (16 v + t) / bs of the row, i.e. byte (16 v + t) / bs - s0 of the 8 samples loaded from s0 = 16 v / bs
    uint32_t sel[2][4];
    int s0v[2];
    #pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int b0 = 16 * (lane + 32 * q);
        s0v[q] = b0 / bs;
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t sj = 0;
            #pragma unroll
            for (int t = 0; t < 4; ++t) sj |= (uint32_t)(((b0 + 4 * j + t) / bs - s0v[q]) & 7) << (4 * t);
            sel[q][j] = sj;
        }
    }
    const int cb = g.chunk;                                      // blocks per chunk: 8, or 2 in small calls (jb_call_chunk_blocks)
    const int nvec = (cb * side) >> 4;                           // 16-byte vectors per pixel row of a chunk
    const bool fast_store_ok = aligned && bs >= 3 && bs <= 8 && a.row_pitch < (1u << 24);     // (32-bit offsets inside a chunk row)

    const unsigned total_warps = gridDim.x * NWARPS;
    // (first chunk dealt statically, CTA-major: a single frame spreads over all SMs; the rest from the counter, whose
    // answer is picked up only at the end of the chunk: the round trip of the atomic hides behind the work)
    unsigned next_chunk = blockIdx.x + gridDim.x * (unsigned)warp;
    while (next_chunk < a.n_chunks) {
        const unsigned chunk = next_chunk;
        unsigned pend = 0;
        if (a.ticket != nullptr && lane == 0) pend = total_warps + atomicAdd(a.ticket, 1u);
        const int plane = (int)(chunk / (unsigned)g.cpp);
        const int blk0 = (int)(chunk % (unsigned)g.cpp) * cb;
        const int nvalid = jb_min(cb, g.nblocks - blk0);

        // ---- coefficients of the chunk in natural order ----
        {
            uint4* z = (uint4*)coef;
            for (int i = lane; i < cb * n / 8; i += 32) z[i] = make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
        int my_umax = D - 1, my_vmax = D - 1;                   // of the block this lane decoded (lane < nvalid)
        if (MODE == 2) {
            for (int gi = 0; gi < nvalid; ++gi)
                for (int zp = lane; zp < n; zp += 32)
                    coef[gi * n + sIzz[zp]] = a.coeffs_in[((size_t)plane * g.nblocks + blk0 + gi) * n + zp];
        } else if (lane < nvalid) {
            const unsigned long long len = a.plane_len[plane];
            const unsigned start = a.block_start[(size_t)plane * g.nblocks + blk0 + lane];
            int rc = 1;
            if (len <= 0xFFFFFFFFull && start < (unsigned)len && a.plane_off[plane] + len <= a.in_bytes)
                rc = il_decode_block<D>(a.in + a.plane_off[plane], start, (unsigned)len, n, coef + lane * n, sIzz, my_umax, my_vmax);
            if (rc) { jb_set_error(a.status, JB_ERR_BAD_STREAM); my_umax = my_vmax = D - 1; }
        }
        __syncwarp();

        for (int gi = 0; gi < nvalid; ++gi) {
            // A heavily quantised block keeps its coefficients in the low-frequency corner (config 3: divide 1000, zigzag
            // positions 0..~10): rows u >= ulim and columns k >= klim (multiples of 8 around the decoder's umax / vmax) hold
            // zeros, so the sums below stop there -- a third of the multiply-adds for an 8 x 8 corner of a 24 x 24 block.
            // Whatever earlier blocks left outside that corner of sY / sP is never read.
            const int ulim = jb_min(D, (__shfl_sync(0xffffffffu, my_umax, gi) + 8) & ~7);
            const int klim = jb_min(D, (__shfl_sync(0xffffffffu, my_vmax, gi) + 8) & ~7);
            const int mlim = klim >> 1, mulim = ulim >> 1;          // in pairs (even, odd): multiples of 4
            // ---- dequantise (integer coefficient * quantiser step: exact in fp32), split by frequency parity ----
            {
                const int16_t* row = coef + gi * n;
                if (lane < klim)
                    for (int u = 0; u < ulim; ++u)
                        sY[(lane & 1) * (D * H) + u * H + (lane >> 1)] = (float)row[u * D + lane] * sDq[u * D + lane];
            }
            __syncwarp();
            // ---- along the rows: Pe / Po[u][c] = sum_m Y[u][2m + par] B[c][2m + par], c < H; lane tile RT rows x KT samples ----
            {
                const int par = kg >> 1, chalf = kg & 1;
                const float* ybase = sY + par * (D * H) + (rg * RT) * H;
                const float* bbase = sBH + par * (H * H) + (chalf * KT) * H;
                float acc[RT][KT];
                #pragma unroll
                for (int r = 0; r < RT; ++r)
                    #pragma unroll
                    for (int q = 0; q < KT; ++q) acc[r][q] = 0.f;
                #pragma unroll
                for (int m = 0; m < H; m += 4) {
                    if (m >= mlim) break;                               // (warp-uniform)
                    float4 y4[RT], b4[KT];
                    #pragma unroll
                    for (int r = 0; r < RT; ++r) y4[r] = *(const float4*)(ybase + r * H + m);
                    #pragma unroll
                    for (int q = 0; q < KT; ++q) b4[q] = *(const float4*)(bbase + q * H + m);
                    #pragma unroll
                    for (int r = 0; r < RT; ++r)
                        #pragma unroll
                        for (int q = 0; q < KT; ++q) {
                            acc[r][q] = fmaf(y4[r].x, b4[q].x, acc[r][q]); acc[r][q] = fmaf(y4[r].y, b4[q].y, acc[r][q]);
                            acc[r][q] = fmaf(y4[r].z, b4[q].z, acc[r][q]); acc[r][q] = fmaf(y4[r].w, b4[q].w, acc[r][q]);
                        }
                }
                // P[u][c] = Pe + Po (the even-parity lane stores it), P[u][D-1-c] = Pe - Po (the odd-parity lane);
                // stored split by the parity of u for the column pass: sP[(u & 1)][u >> 1][column]
                #pragma unroll
                for (int r = 0; r < RT; ++r) {
                    const int u = rg * RT + r;
                    float* prow = sP + (u & 1) * (H * D) + (u >> 1) * D;
                    #pragma unroll
                    for (int q = 0; q < KT; ++q) {
                        const float other = __shfl_xor_sync(0xffffffffu, acc[r][q], 2);
                        const int c = chalf * KT + q;
                        if (par == 0) prow[c] = acc[r][q] + other;
                        else prow[D - 1 - c] = other - acc[r][q];
                    }
                }
            }
            __syncwarp();
            // ---- down the columns: Xe / Xo[r][c] = sum_mu B[r][2 mu + par] P[2 mu + par][c], r < H; lane tile RT sample rows
            //      (one parity of u) x KT columns; np.round, clamp, 8-bit samples ----
            {
                const int par = rg >> 2, rgl = rg & 3;
                const float* bbase = sBH + par * (H * H) + (rgl * RT) * H;
                const float* pbase = sP + par * (H * D) + kg * KT;
                float acc[RT][KT];
                #pragma unroll
                for (int r = 0; r < RT; ++r)
                    #pragma unroll
                    for (int q = 0; q < KT; ++q) acc[r][q] = 0.f;
                #pragma unroll
                for (int mu = 0; mu < H; mu += 4) {
                    if (mu >= mulim) break;                             // (warp-uniform)
                    float4 b4[RT];
                    #pragma unroll
                    for (int r = 0; r < RT; ++r) b4[r] = *(const float4*)(bbase + r * H + mu);
                    #pragma unroll
                    for (int ii = 0; ii < 4; ++ii) {
                        float pv[KT];
                        #pragma unroll
                        for (int q = 0; q < KT; q += 2) {
                            const float2 v2 = *(const float2*)(pbase + (mu + ii) * D + q);
                            pv[q] = v2.x; pv[q + 1] = v2.y;
                        }
                        #pragma unroll
                        for (int r = 0; r < RT; ++r) {
                            const float bv = ii == 0 ? b4[r].x : ii == 1 ? b4[r].y : ii == 2 ? b4[r].z : b4[r].w;
                            #pragma unroll
                            for (int q = 0; q < KT; ++q) acc[r][q] = fmaf(bv, pv[q], acc[r][q]);
                        }
                    }
                }
                #pragma unroll
                for (int r = 0; r < RT; ++r) {
                    const int rr = rgl * RT + r;
                    const int srow = par == 0 ? rr : D - 1 - rr;
                    uint8_t* sp = samp + srow * ROWB + gi * D + kg * KT;
                    #pragma unroll
                    for (int q = 0; q < KT; ++q) {
                        const float other = __shfl_xor_sync(0xffffffffu, acc[r][q], 16);
                        const float x = par == 0 ? acc[r][q] + other : other - acc[r][q];
                        const int pq = __float_as_int(x + 12582912.0f) - 0x4B400000;      // round half-even
                        sp[q] = (uint8_t)max(0, min(255, pq));
                    }
                }
            }
            __syncwarp();
        }

        // ---- the chunk's pixels ----
        uint8_t* dstp = plane0 + (size_t)plane * a.plane_stride;
        const int by0 = blk0 / hb, bx0 = blk0 - by0 * hb;
        const int x0 = bx0 * side, y0 = by0 * side;
        const bool one_row = nvalid == cb && bx0 + cb <= hb &&
                             x0 + cb * side <= g.W && y0 + side <= g.H && (x0 & 15) == 0 && ((cb * side) & 15) == 0;
        if (fast_store_ok && one_row) {
            uint8_t* base = dstp + (size_t)y0 * a.row_pitch + x0;
            const unsigned pitch32 = (unsigned)a.row_pitch;
            #pragma unroll 1
            for (int i = 0; i < D; ++i) {
                const uint8_t* srow = samp + i * ROWB;
                #pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int v = lane + 32 * q;
                    if (v < nvec) {
                        const int s0 = s0v[q];
                        const uint32_t* w = (const uint32_t*)(srow + (s0 & ~3));
                        const uint32_t sh = (uint32_t)(s0 & 3) * 8u;
                        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
                        const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
                        const uint4 o = make_uint4(__byte_perm(lo, hi, sel[q][0]), __byte_perm(lo, hi, sel[q][1]),
                                                   __byte_perm(lo, hi, sel[q][2]), __byte_perm(lo, hi, sel[q][3]));
                        // (offsets inside one chunk row of blocks fit 32 bits: side <= 128 rows of one plane row each)
                        uint8_t* d = base + (size_t)((unsigned)(i * bs) * pitch32 + 16u * (unsigned)v);
                        if (BS) {
                            #pragma unroll
                            for (int k = 0; k < BS; ++k) *(uint4*)(d + (unsigned)k * pitch32) = o;
                        } else {
                            for (int k = 0; k < bs; ++k) *(uint4*)(d + (unsigned)k * pitch32) = o;
                        }
                    }
                }
            }
        } else {
            // edge / wrapped / unaligned chunk: byte stores, cropped to the subsampled extent and to the plane
            for (int gi = 0; gi < nvalid; ++gi) {
                const int blk = blk0 + gi;
                const int by = blk / hb, bx = blk - by * hb;
                const int yy0 = by * side, xx0 = bx * side;
                const int rows = jb_min(side, jb_min(g.H, g.H1 * bs) - yy0), cols = jb_min(side, jb_min(g.W, g.W1 * bs) - xx0);
                if (rows <= 0 || cols <= 0) continue;
                for (int idx = lane; idx < rows * cols; idx += 32) {
                    const int r = idx / cols, cc = idx - r * cols;
                    dstp[(size_t)(yy0 + r) * a.row_pitch + xx0 + cc] = samp[(r / bs) * ROWB + gi * D + cc / bs];
                }
            }
        }
        __syncwarp();                                            // the sample buffer is rewritten by the next chunk
        next_chunk = a.ticket != nullptr ? __shfl_sync(0xffffffffu, pend, 0) : chunk + total_warps;
    }
    jb_dec_epilogue(a);
}

template <int D, int MODE, int BS>
static cudaError_t il_launch_t(const JbInvArgs& a, cudaStream_t s) {
    const IlLayout L = il_layout(D);
    cudaError_t e = cudaFuncSetAttribute(jb_inv_large_kernel<D, MODE, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned want = a.n_chunks;
    const unsigned grid = want < (unsigned)sms ? want : (unsigned)sms;
    if (grid == 0) return cudaSuccess;
    return jb_launch_ex(jb_inv_large_kernel<D, MODE, BS>, dim3(grid), dim3(L.warps * 32), L.total, s,
                        (a.g.flags & JB_FLAG_PDL) != 0, a);
}

template <int MODE>
static cudaError_t il_launch_d(const JbInvArgs& a, cudaStream_t s) {
    switch (a.g.d) {
    case 16: return il_launch_t<16, MODE, 0>(a, s);
    case 24: return a.g.bs == 5 ? il_launch_t<24, MODE, 5>(a, s) : il_launch_t<24, MODE, 0>(a, s);
    default: return il_launch_t<32, MODE, 0>(a, s);
    }
}

cudaError_t jb_launch_inv_large(const JbInvArgs& a, int mode, cudaStream_t s) {
    if (a.n_chunks == 0) return cudaSuccess;
    return mode == 0 ? il_launch_d<0>(a, s) : il_launch_d<2>(a, s);
}
