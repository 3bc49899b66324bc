// jb_forward_mid.cu -- compress direction for "large block" configurations: dct_size a multiple of 4
// (8..32), any block_size whose source tile (d*bs)^2 fits shared memory -- BASELINE.json config 3
// (--block_size 5 --dct_size 24).  One CTA per chunk of 32 blocks, the blocks go through the CTA one
// after another:
//   1. the (d*bs) x (d*bs) source tile is copied to shared memory with 8-byte coalesced loads
//      (edge / misaligned tiles: clamped byte loads -- padding.py:8-12, dct_padding.py:8-9);
//   2. separable box sums: vertical sums of bs rows (32-bit loads, two 16-bit partial sums per register),
//      then horizontal sums of bs columns -> exact integer sums X (subsampling.py:9-11 without the division);
//   3. C.X.C^T as two register-tiled contractions out of shared memory (transforms.py:46-58):
//      each thread owns 4 outputs and streams X and 4-wide slices of C^T -- 2 shared loads per 4 FFMA.
//      fp32 on exact integers; coefficients within the fp32 error bound of a rounding tie are
//      re-evaluated in fp64 (jb_tables.cu);
//   4. quantise, zigzag, int16 rows; a non-zero bitmap per block is built with ballots so that
//   5. the run-length / bit-packing stage (one lane per block, util.py:146-160,203-221) visits only
//      the non-zero coefficients;
//   6. the chunk's bytes go, compacted, to its slot; jb_launch_scan_gather places them.
// A tensor-core variant was considered and rejected: the contraction is ~2 MAC per source byte, far
// below the FFMA roof at HBM speed, and tf32/bf16 cannot carry a 576-term sum of 8-bit data to
// rounding accuracy.
#include "jb_common.cuh"
#include "jb_forward.cuh"
#include "jb_refine.cuh"

#define FM_THREADS 256
#define FM_STAGE_BYTES 128            // staging row per block; longer blocks are packed again, bytes to the slot
#define FM_BIG_CAP 16

struct FmLayout {
    int side;            // d * bs: rows and bytes per row of the source tile
    int pitch;           // bytes per tile row in shared memory (multiple of 8)
    int coefW;           // words per coefficient row (odd)
    int maskW;           // words of non-zero bitmap per block
    int stageW;          // words per staging row (odd)
    size_t at, bt, qm, qt, zz, tile, hsum, x, t, t2, coef, mask, stage, total;
};

__host__ __device__ inline FmLayout fm_layout(int d, int bs, bool dft) {
    FmLayout L;
    const int n = d * d;
    L.side = d * bs;
    L.pitch = (L.side + 7) / 8 * 8;
    L.coefW = ((n + 1) / 2) | 1;
    L.maskW = (n + 31) / 32;
    L.stageW = (FM_STAGE_BYTES / 4) | 1;
    size_t o = 0;
    L.at = o;    o += (size_t)n * 4;
    L.bt = o;    o += dft ? (size_t)n * 4 : 0;
    L.qm = o;    o += (size_t)n * 4;
    L.qt = o;    o += (size_t)n * 4;
    L.zz = o;    o += jb_align_up((size_t)n * 2, 16);
    L.tile = o;  o += 2 * jb_align_up((size_t)L.side * L.pitch, 16);      // double buffered
    L.hsum = o;  o += jb_align_up((size_t)L.pitch * d * 2, 16);      // vertical sums V[i][x], uint16
    L.x = o;     o += (size_t)n * 4;
    L.t = o;     o += (size_t)n * 4;
    L.t2 = o;    o += dft ? (size_t)n * 4 : 0;
    L.coef = o;  o += (size_t)JB_CHUNK * L.coefW * 4;
    L.mask = o;  o += (size_t)JB_CHUNK * L.maskW * 4;
    L.stage = o; o += (size_t)JB_CHUNK * L.stageW * 4;
    L.total = o;
    return L;
}

bool jb_fwd_mid_eligible(const JbGeom& g) {
    if (g.d < 8 || g.d % 4 != 0 || g.d > JB_MAX_DCT_SIZE) return false;
    return fm_layout(g.d, g.bs, g.transform == JB_TRANSFORM_DFT).total <= 200 * 1024;
}

// fp64 re-evaluation from float box sums (jb_refine.cuh), out of line
__device__ __noinline__ double fm_refine(const float* X, int u, int v, int d, int bs, int transform, int qmode,
                                         const double* A64, const double* B64, double recip) {
    return jb_refine_f64<float>(X, u, v, d, bs, transform, qmode, A64, B64, recip);
}

template <typename Writer>
__device__ __forceinline__ void fm_pack_masked(const int16_t* c, const uint32_t* mask, int mask_words, Writer& bw,
                                               int& bad_pos, int& bad_run) {
    int prev = -1;
    bad_pos = -1; bad_run = 0;
    for (int wi = 0; wi < mask_words; ++wi) {
        uint32_t m = mask[wi];
        while (m) {
            const int p = wi * 32 + __ffs((int)m) - 1;
            m &= m - 1;
            const int amp = c[p];
            const int run = p - prev - 1;
            prev = p;
            if (!jb_put_coefficient(bw, run, amp) && bad_pos < 0) { bad_pos = p; bad_run = run % JB_MAX_RUN; }
        }
    }
    bw.put(0u, 8);
}

template <bool DFT, int MODE>
__global__ void __launch_bounds__(FM_THREADS)
jb_fwd_mid_kernel(const JbFwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const JbGeom& g = a.g;
    const int d = g.d, n = g.n, bs = g.bs, q4 = d >> 2;
    const FmLayout L = fm_layout(d, bs, DFT);
    float* sAt = (float*)(smem + L.at);          // At[j][v] = A[v][j]
    float* sBt = (float*)(smem + L.bt);
    float* sQm = (float*)(smem + L.qm);
    float* sQt = (float*)(smem + L.qt);
    uint16_t* sZz = (uint16_t*)(smem + L.zz);
    uint8_t* sTile = smem + L.tile;
    uint16_t* sH = (uint16_t*)(smem + L.hsum);
    float* sX = (float*)(smem + L.x);
    float* sT = (float*)(smem + L.t);
    float* sT2 = (float*)(smem + L.t2);
    uint32_t* sCoef = (uint32_t*)(smem + L.coef);
    uint32_t* sMask = (uint32_t*)(smem + L.mask);
    uint32_t* sStage = (uint32_t*)(smem + L.stage);

    __shared__ unsigned s_chunk;
    __shared__ unsigned s_blen[JB_CHUNK], s_boff[JB_CHUNK];
    __shared__ int s_big_blk[FM_BIG_CAP], s_big_pos[FM_BIG_CAP], s_big_amp[FM_BIG_CAP];
    __shared__ int s_nbig, s_slow, s_P;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        s_P = jb_ctrl_parity(a);
        s_chunk = atomicAdd(jb_ctrl_ticket(a, s_P), 1u);
        jb_ctrl_note_first(a, s_P, s_chunk);
        s_nbig = 0; s_slow = 0;
    }
    for (int idx = tid; idx < n; idx += FM_THREADS) {
        const int v = idx / d, j = idx - v * d;
        sAt[j * d + v] = a.t.fA[idx];
        if (DFT) sBt[j * d + v] = a.t.fB[idx];
        sQm[idx] = a.t.qmult[idx];
        sQt[idx] = a.t.qtol[idx];
        sZz[idx] = a.t.zz[idx];
    }
    __syncthreads();
    const unsigned chunk = s_chunk;
    const int P = s_P;
    if (chunk >= a.n_chunks) return;
    const int plane = chunk / g.cpp;
    const int blk0 = (chunk % g.cpp) * JB_CHUNK;
    const int nvalid = jb_min(JB_CHUNK, g.nblocks - blk0);
    const uint8_t* src = a.planes + (size_t)plane * a.plane_stride;
    const int side = L.side, pitch = L.pitch;
    const bool vec_ok = (side % 8 == 0) && (((uintptr_t)src & 7) == 0) && (a.row_pitch % 8 == 0);
    const bool refine_on = !(g.flags & JB_FLAG_NO_REFINE);

    // thread roles of the two contractions: (row/col index, group of 4 outputs)
    const int ci = tid / q4, cq = tid - ci * q4;
    const bool c_live = ci < d;

    // ---- 1. source tile -> shared memory, one block ahead of the arithmetic (cp.async) ----
    const size_t tile_bytes = jb_align_up((size_t)side * pitch, 16);
    const int w8 = jb_min(side >> 3, FM_THREADS) > 0 ? (side >> 3) : 1;
    const int vr0 = tid / w8, vc0 = tid - vr0 * w8, vdr = FM_THREADS / w8, vdc = FM_THREADS - vdr * w8;
    auto stage_tile = [&](int gi, uint8_t* buf) {
        const int blk = blk0 + gi;
        const int by = blk / g.hb, bx = blk - by * g.hb;
        const bool interior = (by + 1) * side <= g.H && (bx + 1) * side <= g.W;
        if (interior && vec_ok) {
            const uint8_t* base = src + (size_t)by * side * a.row_pitch + (size_t)bx * side;
            int r = vr0, c = vc0;                       // (row, 8-byte word) of this thread, stepped without divisions
            while (r < side) {
                const unsigned sdst = (unsigned)__cvta_generic_to_shared(buf + r * pitch + c * 8);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;"
                             :: "r"(sdst), "l"(base + (size_t)r * a.row_pitch + (size_t)c * 8) : "memory");
                r += vdr; c += vdc;
                if (c >= w8) { c -= w8; ++r; }
            }
        } else {
            for (int idx = tid; idx < side * side; idx += FM_THREADS) {
                const int r = idx / side, c = idx - r * side;
                const int si = jb_min(by * d + r / bs, g.H1 - 1), sj = jb_min(bx * d + c / bs, g.W1 - 1);
                const int y = jb_min(si * bs + r % bs, g.H - 1), x = jb_min(sj * bs + c % bs, g.W - 1);
                buf[r * pitch + c] = src[(size_t)y * a.row_pitch + x];
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage_tile(0, sTile);

    for (int gi = 0; gi < nvalid; ++gi) {
        const int blk = blk0 + gi;
        uint8_t* tile = sTile + (size_t)(gi & 1) * tile_bytes;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                       // tile gi complete; everyone is done with block gi - 1
        if (gi + 1 < nvalid) stage_tile(gi + 1, sTile + (size_t)((gi + 1) & 1) * tile_bytes);
        // ---- 2a. vertical sums of bs rows, four byte columns per thread (two 16-bit lanes per register) ----
        {
            const int nwc = pitch >> 2;                                   // words per tile row
            const uint32_t* t32 = (const uint32_t*)tile;
            uint32_t* v32 = (uint32_t*)sH;                                // V[i][x] as uint16, row length = pitch
            for (int t = tid; t < d * nwc; t += FM_THREADS) {
                const int i = t / nwc, wc = t - i * nwc;
                const uint32_t* p = t32 + (size_t)i * bs * nwc + wc;
                uint32_t even = 0, odd = 0;                               // bytes 0,2 and bytes 1,3 (sums <= 255 bs < 2^16)
                for (int k = 0; k < bs; ++k) {
                    const uint32_t w = p[k * nwc];
                    even += w & 0x00FF00FFu;
                    odd += (w >> 8) & 0x00FF00FFu;
                }
                v32[i * (nwc * 2) + 2 * wc] = (even & 0xFFFFu) | (odd << 16);
                v32[i * (nwc * 2) + 2 * wc + 1] = (even >> 16) | (odd & 0xFFFF0000u);
            }
        }
        __syncthreads();
        // ---- 2b. horizontal sums of bs columns -> X (exact integers, as float) ----
        for (int t = tid; t < n; t += FM_THREADS) {
            const int i = t / d, j = t - i * d;
            const uint16_t* p = sH + i * pitch + j * bs;
            int sum = 0;
            for (int k = 0; k < bs; ++k) sum += p[k];
            sX[t] = (float)sum;
        }
        __syncthreads();
        // ---- 3a. T[i][v] = sum_j X[i][j] A[v][j]: thread (i, group of 4 v) ----
        if (c_live) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), acc2 = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* xr = sX + ci * d;
            for (int j = 0; j < d; ++j) {
                const float x = xr[j];
                const float4 c4 = *(const float4*)(sAt + j * d + 4 * cq);
                acc.x = fmaf(x, c4.x, acc.x); acc.y = fmaf(x, c4.y, acc.y);
                acc.z = fmaf(x, c4.z, acc.z); acc.w = fmaf(x, c4.w, acc.w);
                if (DFT) {
                    const float4 s4 = *(const float4*)(sBt + j * d + 4 * cq);
                    acc2.x = fmaf(x, s4.x, acc2.x); acc2.y = fmaf(x, s4.y, acc2.y);
                    acc2.z = fmaf(x, s4.z, acc2.z); acc2.w = fmaf(x, s4.w, acc2.w);
                }
            }
            *(float4*)(sT + ci * d + 4 * cq) = acc;
            if (DFT) *(float4*)(sT2 + ci * d + 4 * cq) = acc2;
        }
        __syncthreads();
        // ---- 3b. Y[u][v] = sum_i A[u][i] T[i][v] (- B[u][i] T2[i][v]): thread (v, group of 4 u) ----
        if (c_live) {
            const int v = ci;
            float y4[4] = {0.f, 0.f, 0.f, 0.f};
            for (int i = 0; i < d; ++i) {
                const float t = sT[i * d + v];
                const float4 c4 = *(const float4*)(sAt + i * d + 4 * cq);
                y4[0] = fmaf(c4.x, t, y4[0]); y4[1] = fmaf(c4.y, t, y4[1]);
                y4[2] = fmaf(c4.z, t, y4[2]); y4[3] = fmaf(c4.w, t, y4[3]);
                if (DFT) {
                    const float t2 = sT2[i * d + v];
                    const float4 s4 = *(const float4*)(sBt + i * d + 4 * cq);
                    y4[0] = fmaf(-s4.x, t2, y4[0]); y4[1] = fmaf(-s4.y, t2, y4[1]);
                    y4[2] = fmaf(-s4.z, t2, y4[2]); y4[3] = fmaf(-s4.w, t2, y4[3]);
                }
            }
            // ---- 4. quantise, tie check, zigzag ----
            #pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int u = 4 * cq + k, idx = u * d + v;
                const float val = y4[k] * sQm[idx];
                float r = rintf(val);
                if (refine_on && fabsf(fabsf(val - r) - 0.5f) < sQt[idx] + 2.4e-7f * fabsf(val))
                    r = (float)rint(fm_refine(sX, u, v, d, bs, g.transform, g.qmode, a.t.fA64, a.t.fB64, a.t.qrecip[idx]));
                int q = (int)r;
                const int zp = sZz[idx];
                if (MODE == 1) {
                    a.coeffs_out[((size_t)plane * g.nblocks + blk) * n + zp] = (int16_t)max(-32767, min(32767, q));
                } else {
                    if (q > JB_MAX_AMP || q < -JB_MAX_AMP) {
                        const int kk = atomicAdd(&s_nbig, 1);
                        if (kk < FM_BIG_CAP) { s_big_blk[kk] = gi; s_big_pos[kk] = zp; s_big_amp[kk] = q; }
                        q = q > 0 ? 32767 : -32767;
                    }
                    ((int16_t*)(sCoef + gi * L.coefW))[zp] = (int16_t)q;
                }
            }
        }
        __syncthreads();
        // ---- non-zero bitmap of the block (zigzag order) ----
        if (MODE != 1) {
            const int16_t* row = (const int16_t*)(sCoef + gi * L.coefW);
            for (int w0 = warp; w0 < L.maskW; w0 += FM_THREADS / 32) {
                const int p = w0 * 32 + lane;
                const unsigned m = __ballot_sync(0xffffffffu, p < n && row[p] != 0);
                if (lane == 0) sMask[gi * L.maskW + w0] = m;
            }
        }
        // (the next block's tile load does not touch anything the bitmap pass reads)
    }
    if (MODE == 1) return;
    __syncthreads();

    // ---- 5. run-length + bit packing, one lane per block, non-zero coefficients only ----
    if (tid < JB_CHUNK) {
        unsigned len = 0;
        if (tid < nvalid) {
            const int16_t* c = (const int16_t*)(sCoef + tid * L.coefW);
            JbBitWriter bw;
            bw.init(sStage + tid * L.stageW, FM_STAGE_BYTES / 4);
            int bad_pos, bad_run;
            fm_pack_masked(c, sMask + tid * L.maskW, L.maskW, bw, bad_pos, bad_run);
            len = bw.finish();
            if (len > FM_STAGE_BYTES) s_slow = 1;
            if (bad_pos >= 0) {
                long long amp = c[bad_pos];
                const int nb = jb_min(s_nbig, FM_BIG_CAP);
                for (int k = 0; k < nb; ++k)
                    if (s_big_blk[k] == tid && s_big_pos[k] == bad_pos) amp = s_big_amp[k];
                jb_report_bad_code(jb_ctrl_status(a, P), (unsigned long long)plane * g.nblocks + blk0 + tid, bad_pos, bad_run, amp);
            }
        }
        unsigned incl = len;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
            if (tid >= o) incl += y;
        }
        s_blen[tid] = len;
        s_boff[tid] = incl - len;
        if (tid == 31) jb_record_chunk_len(a, P, chunk, incl);
    }
    __syncthreads();
    // ---- 6. chunk bytes -> slot ----
    uint8_t* slot = jb_chunk_slot(a, chunk, s_boff[JB_CHUNK - 1] + s_blen[JB_CHUNK - 1]);
    if (!s_slow) {
        for (int gi = 0; gi < nvalid; ++gi) {
            const uint8_t* sb = (const uint8_t*)(sStage + gi * L.stageW);
            uint8_t* dst = slot + s_boff[gi];
            for (unsigned j = tid; j < s_blen[gi]; j += FM_THREADS) dst[j] = sb[j];
        }
    } else if (tid < nvalid) {
        const int16_t* c = (const int16_t*)(sCoef + tid * L.coefW);
        JbByteWriter bw;
        bw.init(slot + s_boff[tid]);
        int bad_pos, bad_run;
        fm_pack_masked(c, sMask + tid * L.maskW, L.maskW, bw, bad_pos, bad_run);
        bw.finish();
    }
}

// =====================================================================================================
// Warp-per-block variant (DCT, dct_size 16 / 24 / 32, source tile rows of at most 128 bytes).
//
// The CTA-per-block kernel above spends its time at barriers: every block takes five CTA-wide phases with
// at most 144 busy threads each (ncu, config 3: 39 % of the issue slots used, 16 % of the stall samples at
// the barrier, the contractions themselves only 15 % of the instructions).  Here every warp takes a block of
// the chunk through all stages on its own, so eight blocks per CTA are in flight and the CTA meets only
// twice per chunk:
//   * box sums straight from global memory: lane c loads word column c of the bs rows of a band (coalesced
//     4-byte loads, several bands in flight), adds them as two 16-bit lanes per register, and bs neighbouring
//     column sums are added through a 256-byte scratch row -> row i of X (exact integers, as float);
//   * C.X.C^T with the operands in registers: lane v keeps row v of C and computes T[i][v] for all i from
//     broadcast reads of X, then Y[u][v] for all u from broadcast reads of C -- T never leaves the registers.
//     The summation order is the one of the kernel above (fp32 fmaf chains over ascending index), so both
//     produce the same coefficients;
//   * quantise, tie check (fp64 re-evaluation), zigzag, int16 row, non-zero bitmap by ballots;
// then, as above, one lane per block run-length encodes and packs, and the CTA copies the chunk to its slot.
// =====================================================================================================
#define FW_WARPS 8
#define FW_BANDS 4                      // bands whose loads are issued together (divides 16, 24, 32)

struct FwLayout {
    int coefW, maskW, stageW, vrowW;
    size_t a, qm, qt, zz, x, vrow, coef, mask, stage, total;
};

__host__ __device__ inline FwLayout fw_layout(int d, int side) {
    FwLayout L;
    const int n = d * d;
    L.coefW = ((n + 1) / 2) | 1;
    L.maskW = (n + 31) / 32;
    L.stageW = (FM_STAGE_BYTES / 4) | 1;
    L.vrowW = (side + 7) / 8 * 4 + 4;                   // uint16 column sums of one band, in words (4 per 32-bit source word)
    size_t o = 0;
    L.a = o;     o += (size_t)n * 4;
    L.qm = o;    o += (size_t)n * 4;
    L.qt = o;    o += (size_t)n * 4;
    L.zz = o;    o += jb_align_up((size_t)n * 2, 16);
    L.x = o;     o += (size_t)FW_WARPS * n * 4;
    L.vrow = o;  o += jb_align_up((size_t)FW_WARPS * 2 * L.vrowW * 4, 16);
    L.coef = o;  o += (size_t)JB_CHUNK * L.coefW * 4;
    L.mask = o;  o += (size_t)JB_CHUNK * L.maskW * 4;
    L.stage = o; o += (size_t)JB_CHUNK * L.stageW * 4;
    L.total = o;
    return L;
}

static bool fw_eligible(const JbGeom& g) {
    if (g.transform != JB_TRANSFORM_DCT) return false;
    if (g.d != 16 && g.d != 24 && g.d != 32) return false;
    const int side = g.d * g.bs;
    return side <= 128 && fw_layout(g.d, side).total <= 110 * 1024;
}

template <int D, int MODE>
__global__ void __launch_bounds__(FW_WARPS * 32, 2)
jb_fwd_mid_warp_kernel(const JbFwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const JbGeom& g = a.g;
    constexpr int n = D * D;
    const int bs = g.bs, side = D * bs;
    const FwLayout L = fw_layout(D, side);
    float* sA = (float*)(smem + L.a);            // A[u][i], row-major
    float* sQm = (float*)(smem + L.qm);
    float* sQt = (float*)(smem + L.qt);
    uint16_t* sZz = (uint16_t*)(smem + L.zz);
    uint32_t* sCoef = (uint32_t*)(smem + L.coef);
    uint32_t* sMask = (uint32_t*)(smem + L.mask);
    uint32_t* sStage = (uint32_t*)(smem + L.stage);

    __shared__ unsigned s_chunk;
    __shared__ unsigned s_blen[JB_CHUNK], s_boff[JB_CHUNK];
    __shared__ int s_big_blk[FM_BIG_CAP], s_big_pos[FM_BIG_CAP], s_big_amp[FM_BIG_CAP];
    __shared__ int s_nbig, s_slow, s_P;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* sX = (float*)(smem + L.x) + warp * n;
    uint16_t* sV = (uint16_t*)(smem + L.vrow) + (size_t)warp * 2 * L.vrowW * 2;     // two band rows per warp
    if (tid == 0) {
        s_P = jb_ctrl_parity(a);
        s_chunk = atomicAdd(jb_ctrl_ticket(a, s_P), 1u);
        jb_ctrl_note_first(a, s_P, s_chunk);
        s_nbig = 0; s_slow = 0;
    }
    for (int idx = tid; idx < n; idx += FW_WARPS * 32) {
        sA[idx] = a.t.fA[idx];
        sQm[idx] = a.t.qmult[idx];
        sQt[idx] = a.t.qtol[idx];
        sZz[idx] = a.t.zz[idx];
    }
    __syncthreads();
    const unsigned chunk = s_chunk;
    const int P = s_P;
    if (chunk >= a.n_chunks) return;
    const int plane = chunk / g.cpp;
    const int blk0 = (chunk % g.cpp) * JB_CHUNK;
    const int nvalid = jb_min(JB_CHUNK, g.nblocks - blk0);
    const uint8_t* src = a.planes + (size_t)plane * a.plane_stride;
    const bool vec_ok = (side % 4 == 0) && (((uintptr_t)src & 3) == 0) && (a.row_pitch % 4 == 0);
    const bool refine_on = !(g.flags & JB_FLAG_NO_REFINE);
    const int nwc = side >> 2;                                  // 32-bit words per tile row (<= 32)
    const int v = lane < D ? lane : D - 1;                      // frequency column of this lane (lanes >= D idle along)

    for (int gi = warp; gi < nvalid; gi += FW_WARPS) {
        const int blk = blk0 + gi;
        const int by = blk / g.hb, bx = blk - by * g.hb;
        const bool interior = (by + 1) * side <= g.H && (bx + 1) * side <= g.W;
        // ---- box sums -> X ----
        if (interior && vec_ok) {
            const uint32_t* base = (const uint32_t*)(src + (size_t)by * side * a.row_pitch + (size_t)bx * side);
            const size_t pw = a.row_pitch >> 2;
            // FW_BANDS bands at a time, and the loads of the next group are issued before this group's sums go through
            // the scratch rows: two groups (2 FW_BANDS bs loads per lane) are in flight
            uint32_t even[FW_BANDS], odd[FW_BANDS], even_n[FW_BANDS], odd_n[FW_BANDS];
            const size_t band = (size_t)bs * pw;
            auto load_group = [&](int i0, uint32_t (&ev)[FW_BANDS], uint32_t (&od)[FW_BANDS]) {
                #pragma unroll
                for (int b = 0; b < FW_BANDS; ++b) ev[b] = od[b] = 0u;
                if (lane < nwc && i0 < D) {
                    const uint32_t* p[FW_BANDS];                    // one row pointer per band, stepped down the bs rows
                    p[0] = base + (size_t)i0 * band + lane;
                    #pragma unroll
                    for (int b = 1; b < FW_BANDS; ++b) p[b] = p[b - 1] + band;
                    #pragma unroll 5
                    for (int k = 0; k < bs; ++k) {
                        #pragma unroll
                        for (int b = 0; b < FW_BANDS; ++b) {
                            const uint32_t w = __ldg(p[b]);
                            p[b] += pw;
                            ev[b] += w & 0x00FF00FFu;               // bytes 0,2 and 1,3 of the word column (sums <= 255 bs < 2^16)
                            od[b] += (w >> 8) & 0x00FF00FFu;
                        }
                    }
                }
            };
            load_group(0, even, odd);
            for (int i0 = 0; i0 < D; i0 += FW_BANDS) {
                load_group(i0 + FW_BANDS, even_n, odd_n);
                #pragma unroll
                for (int b = 0; b < FW_BANDS; ++b) {
                    uint16_t* vr = sV + (b & 1) * (L.vrowW * 2);
                    if (lane < nwc)
                        *(uint2*)(vr + 4 * lane) = make_uint2((even[b] & 0xFFFFu) | (odd[b] << 16), (even[b] >> 16) | (odd[b] & 0xFFFF0000u));
                    __syncwarp();
                    if (lane < D) {
                        int sum = 0;
                        for (int k = 0; k < bs; ++k) sum += vr[lane * bs + k];
                        sX[(i0 + b) * D + lane] = (float)sum;
                    }
                    // (the other scratch row serves the next band; this one again two bands later, after two syncs)
                }
                #pragma unroll
                for (int b = 0; b < FW_BANDS; ++b) { even[b] = even_n[b]; odd[b] = odd_n[b]; }
            }
        } else {
            // edge block or unaligned plane: the reference's two-level edge replication, byte by byte
            if (lane < D) {
                const int sj = jb_min(bx * D + lane, g.W1 - 1);
                for (int i = 0; i < D; ++i) {
                    const int si = jb_min(by * D + i, g.H1 - 1);
                    int sum = 0;
                    for (int di = 0; di < bs; ++di) {
                        const uint8_t* r = src + (size_t)jb_min(si * bs + di, g.H - 1) * a.row_pitch;
                        for (int dj = 0; dj < bs; ++dj) sum += r[jb_min(sj * bs + dj, g.W - 1)];
                    }
                    sX[i * D + lane] = (float)sum;
                }
            }
        }
        __syncwarp();
        // ---- T[i][v] = sum_j X[i][j] A[v][j], all i, in registers ----
        float arow[D];                                          // row v of the transform matrix (not live during the loads above)
        #pragma unroll
        for (int j = 0; j < D; ++j) arow[j] = sA[v * D + j];
        float t[D];
        #pragma unroll
        for (int i = 0; i < D; ++i) {
            float acc = 0.f;
            #pragma unroll
            for (int j = 0; j < D; j += 4) {
                const float4 x4 = *(const float4*)(sX + i * D + j);     // same address in every lane: broadcast
                acc = fmaf(x4.x, arow[j], acc); acc = fmaf(x4.y, arow[j + 1], acc);
                acc = fmaf(x4.z, arow[j + 2], acc); acc = fmaf(x4.w, arow[j + 3], acc);
            }
            t[i] = acc;
        }
        // ---- Y[u][v] = sum_i A[u][i] T[i][v]; quantise, tie check, zigzag ----
        int16_t* crow = (int16_t*)(sCoef + gi * L.coefW);
        #pragma unroll 4
        for (int u = 0; u < D; ++u) {
            float y = 0.f;
            #pragma unroll
            for (int i = 0; i < D; i += 4) {
                const float4 c4 = *(const float4*)(sA + u * D + i);     // broadcast
                y = fmaf(c4.x, t[i], y); y = fmaf(c4.y, t[i + 1], y);
                y = fmaf(c4.z, t[i + 2], y); y = fmaf(c4.w, t[i + 3], y);
            }
            if (lane < D) {
                const int idx = u * D + v;
                const float val = y * sQm[idx];
                float r = rintf(val);
                if (refine_on && fabsf(fabsf(val - r) - 0.5f) < sQt[idx] + 2.4e-7f * fabsf(val))
                    r = (float)rint(fm_refine(sX, u, v, D, bs, g.transform, g.qmode, a.t.fA64, a.t.fB64, a.t.qrecip[idx]));
                int q = (int)r;
                const int zp = sZz[idx];
                if (MODE == 1) {
                    a.coeffs_out[((size_t)plane * g.nblocks + blk) * n + zp] = (int16_t)max(-32767, min(32767, q));
                } else {
                    if (q > JB_MAX_AMP || q < -JB_MAX_AMP) {
                        const int kk = atomicAdd(&s_nbig, 1);
                        if (kk < FM_BIG_CAP) { s_big_blk[kk] = gi; s_big_pos[kk] = zp; s_big_amp[kk] = q; }
                        q = q > 0 ? 32767 : -32767;
                    }
                    crow[zp] = (int16_t)q;
                }
            }
        }
        __syncwarp();
        // ---- non-zero bitmap of the block (zigzag order) ----
        if (MODE != 1) {
            for (int w0 = 0; w0 < L.maskW; ++w0) {
                const int p = w0 * 32 + lane;
                const unsigned m = __ballot_sync(0xffffffffu, p < n && crow[p] != 0);
                if (lane == 0) sMask[gi * L.maskW + w0] = m;
            }
        }
        __syncwarp();
    }
    if (MODE == 1) return;
    // ---- run-length + bit packing of the warp's own blocks, one lane per block, non-zero coefficients only:
    //      the eight warps pack side by side instead of queueing behind one packing warp at a barrier ----
    __syncwarp();
    {
        const int my_gi = warp + lane * FW_WARPS;
        if (lane < JB_CHUNK / FW_WARPS) {
            unsigned len = 0;
            if (my_gi < nvalid) {
                const int16_t* c = (const int16_t*)(sCoef + my_gi * L.coefW);
                JbBitWriter bw;
                bw.init(sStage + my_gi * L.stageW, FM_STAGE_BYTES / 4);
                int bad_pos, bad_run;
                fm_pack_masked(c, sMask + my_gi * L.maskW, L.maskW, bw, bad_pos, bad_run);
                len = bw.finish();
                if (len > FM_STAGE_BYTES) s_slow = 1;
                if (bad_pos >= 0) {
                    long long amp = c[bad_pos];
                    const int nb = jb_min(s_nbig, FM_BIG_CAP);
                    for (int k = 0; k < nb; ++k)
                        if (s_big_blk[k] == my_gi && s_big_pos[k] == bad_pos) amp = s_big_amp[k];
                    jb_report_bad_code(jb_ctrl_status(a, P), (unsigned long long)plane * g.nblocks + blk0 + my_gi, bad_pos, bad_run, amp);
                }
            }
            s_blen[my_gi] = len;
        }
    }
    __syncthreads();
    if (tid < JB_CHUNK) {
        const unsigned len = s_blen[tid];
        unsigned incl = len;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
            if (tid >= o) incl += y;
        }
        s_boff[tid] = incl - len;
        if (tid == 31) jb_record_chunk_len(a, P, chunk, incl);
    }
    __syncthreads();
    uint8_t* slot = jb_chunk_slot(a, chunk, s_boff[JB_CHUNK - 1] + s_blen[JB_CHUNK - 1]);
    if (!s_slow) {
        for (int gi = 0; gi < nvalid; ++gi) {
            const uint8_t* sb = (const uint8_t*)(sStage + gi * L.stageW);
            uint8_t* dst = slot + s_boff[gi];
            for (unsigned j = tid; j < s_blen[gi]; j += FW_WARPS * 32) dst[j] = sb[j];
        }
    } else if (tid < nvalid) {
        const int16_t* c = (const int16_t*)(sCoef + tid * L.coefW);
        JbByteWriter bw;
        bw.init(slot + s_boff[tid]);
        int bad_pos, bad_run;
        fm_pack_masked(c, sMask + tid * L.maskW, L.maskW, bw, bad_pos, bad_run);
        bw.finish();
    }
}

template <int D, int MODE>
static cudaError_t fw_launch_t(const JbFwdArgs& a, cudaStream_t s) {
    const size_t smem = fw_layout(D, D * a.g.bs).total;
    cudaError_t e = cudaFuncSetAttribute(jb_fwd_mid_warp_kernel<D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    jb_fwd_mid_warp_kernel<D, MODE><<<a.n_chunks, FW_WARPS * 32, smem, s>>>(a);
    return cudaGetLastError();
}

template <int MODE>
static cudaError_t fw_launch_d(const JbFwdArgs& a, cudaStream_t s) {
    switch (a.g.d) {
    case 16: return fw_launch_t<16, MODE>(a, s);
    case 24: return fw_launch_t<24, MODE>(a, s);
    default: return fw_launch_t<32, MODE>(a, s);
    }
}

template <bool DFT, int MODE>
static cudaError_t fm_launch_t(const JbFwdArgs& a, cudaStream_t s) {
    const size_t smem = fm_layout(a.g.d, a.g.bs, DFT).total;
    cudaError_t e = cudaFuncSetAttribute(jb_fwd_mid_kernel<DFT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    jb_fwd_mid_kernel<DFT, MODE><<<a.n_chunks, FM_THREADS, smem, s>>>(a);
    return cudaGetLastError();
}

cudaError_t jb_launch_fwd_mid(const JbFwdArgs& a, int mode, cudaStream_t s) {
    if (a.n_chunks == 0) return cudaSuccess;
    const bool dft = a.g.transform == JB_TRANSFORM_DFT;
    if (fw_eligible(a.g) && !(a.g.flags & JB_FLAG_NO_TMA)) {       // (JB_FLAG_NO_TMA doubles as "plain variant" for the tests)
        if (mode == 0) {
            cudaError_t e = fw_launch_d<0>(a, s);
            if (e != cudaSuccess) return e;
            return jb_launch_gather(a, s);
        }
        return fw_launch_d<1>(a, s);
    }
    if (mode == 0) {
        cudaError_t e = dft ? fm_launch_t<true, 0>(a, s) : fm_launch_t<false, 0>(a, s);
        if (e != cudaSuccess) return e;
        return jb_launch_gather(a, s);
    }
    return dft ? fm_launch_t<true, 1>(a, s) : fm_launch_t<false, 1>(a, s);
}
