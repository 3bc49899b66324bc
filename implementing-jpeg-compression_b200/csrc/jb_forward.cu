// jb_forward.cu -- compress direction.
//
//   jb_fwd_generic_kernel : any (block_size, dct_size <= 32, DCT | DFT, quantiser); one CTA per
//                           chunk of 32 blocks.  Reference stages 0-8 fused:
//                           padding.py:8-12, subsampling.py:9-11, dct_padding.py:8-9,
//                           normalization.py:7-8, basis_change.py:11-26 (transforms.py:36-58),
//                           quantization.py:8-18 (quantizers.py), zigzag_order.py:85-99,
//                           run_length_encoding.py:47-62, rle_byte_stream.py:48-58.
//   jb_fwd_fast kernels   : see jb_forward_fast.cuh (dct_size 8, block_size 4).
//
// Output order is the reference's: blocks in raster order inside a plane
// (run_length_encoding.py:56-60), planes concatenated.  The transform kernels leave every
// chunk's packed bytes in a per-chunk slot; a device-wide exclusive scan of the chunk lengths
// and a gather pass put them in place (the streams are ~2 % of the pixel bytes, so the extra
// pass costs ~4 % more traffic and needs no inter-block waiting).
#include "jb_common.cuh"
#include "jb_forward.cuh"
#include "jb_refine.cuh"

// ----------------------------------------------------------------------------------------------
// generic kernel
// ----------------------------------------------------------------------------------------------
struct JbBigAmp { int blk, pos, amp; };
#define JB_BIGAMP_CAP 16

struct JbFwdSmemLayout {
    int slot_threads, nsub;
    int coefW;      // 32-bit words per coefficient row (int16 pairs), odd
    int stageW;     // 32-bit words per staging row, odd
    size_t off_A, off_B, off_X, off_T, off_T2, off_coef, off_stage, total;
};

__host__ __device__ inline JbFwdSmemLayout jb_fwd_smem_layout(int d, bool dft) {
    JbFwdSmemLayout L;
    int n = d * d;
    int st = (n + 31) / 32 * 32;
    if (st > JB_GENERIC_THREADS) st = JB_GENERIC_THREADS;
    L.slot_threads = st;
    L.nsub = JB_GENERIC_THREADS / st;
    L.coefW = ((n + 1) / 2) | 1;
    L.stageW = ((jb_max_block_bytes(n) + 3) / 4 + 1) | 1;
    size_t o = 0;
    L.off_A = o;     o += (size_t)n * 4;
    L.off_B = o;     o += dft ? (size_t)n * 4 : 0;
    L.off_X = o;     o += (size_t)L.nsub * n * 4;
    L.off_T = o;     o += (size_t)L.nsub * n * 4;
    L.off_T2 = o;    o += dft ? (size_t)L.nsub * n * 4 : 0;
    L.off_coef = o;  o += (size_t)JB_CHUNK * L.coefW * 4;
    L.off_stage = o; o += (size_t)JB_CHUNK * L.stageW * 4;
    L.total = o;
    return L;
}

size_t jb_fwd_generic_smem_bytes(int d, bool dft) { return jb_fwd_smem_layout(d, dft).total; }

// float64 re-evaluation of one coefficient the way the reference computes it (jb_refine.cuh), out of line
__device__ __noinline__ double jb_refine_coefficient(const int* X, int u, int v, const JbGeom& g, const JbTables& t) {
    return jb_refine_f64<int>(X, u, v, g.d, g.bs, g.transform, g.qmode, t.fA64, t.fB64, t.qrecip[u * g.d + v]);
}

template <int MODE>
__global__ void __launch_bounds__(JB_GENERIC_THREADS)
jb_fwd_generic_kernel(const JbFwdArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const JbGeom& g = a.g;
    const int d = g.d, n = g.n;
    const bool dft = g.transform == JB_TRANSFORM_DFT;
    const JbFwdSmemLayout L = jb_fwd_smem_layout(d, dft);
    float* sA = (float*)(smem + L.off_A);
    float* sB = (float*)(smem + L.off_B);
    int* sX = (int*)(smem + L.off_X);
    float* sT = (float*)(smem + L.off_T);
    float* sT2 = (float*)(smem + L.off_T2);
    uint32_t* sCoef = (uint32_t*)(smem + L.off_coef);
    uint32_t* sStage = (uint32_t*)(smem + L.off_stage);

    __shared__ unsigned s_chunk;
    __shared__ unsigned s_blen[JB_CHUNK], s_boff[JB_CHUNK];
    __shared__ JbBigAmp s_big[JB_BIGAMP_CAP];
    __shared__ int s_nbig, s_P;

    const int tid = threadIdx.x;
    if (tid == 0) {
        s_P = jb_ctrl_parity(a);
        s_chunk = atomicAdd(jb_ctrl_ticket(a, s_P), 1u);
        jb_ctrl_note_first(a, s_P, s_chunk);
        s_nbig = 0;
    }
    if (MODE != 2) {
        for (int i = tid; i < n; i += JB_GENERIC_THREADS) {
            sA[i] = a.t.fA[i];
            if (dft) sB[i] = a.t.fB[i];
        }
    }
    __syncthreads();
    const unsigned chunk = s_chunk;
    const int P = s_P;
    if (chunk >= a.n_chunks) return;
    const int plane = chunk / g.cpp;
    const int blk0 = (chunk % g.cpp) * JB_CHUNK;
    const int nvalid = jb_min(JB_CHUNK, g.nblocks - blk0);

    if (MODE != 2) {
        const uint8_t* src = a.planes + (size_t)plane * a.plane_stride;
        const int slot = tid / L.slot_threads, within = tid % L.slot_threads;
        for (int base = 0; base < nvalid; base += L.nsub) {
            const int gi = base + slot;
            const bool live = slot < L.nsub && gi < nvalid;
            int* X = sX + slot * n;
            float* T = sT + slot * n;
            float* T2 = sT2 + slot * n;
            const int blk = blk0 + gi;
            const int by = blk / g.hb, bx = blk % g.hb;
            // A1-A3: box sums with two-level edge replication
            if (live) {
                for (int idx = within; idx < n; idx += L.slot_threads) {
                    int i = idx / d, j = idx % d;
                    int si = jb_min(by * d + i, g.H1 - 1), sj = jb_min(bx * d + j, g.W1 - 1);
                    int s = 0;
                    for (int di = 0; di < g.bs; ++di) {
                        const uint8_t* row = src + (size_t)jb_min(si * g.bs + di, g.H - 1) * a.row_pitch;
                        for (int dj = 0; dj < g.bs; ++dj) s += row[jb_min(sj * g.bs + dj, g.W - 1)];
                    }
                    X[idx] = s;
                }
            }
            __syncthreads();
            // A5 first pass: T[i][v] = sum_j X[i][j] A[v][j]
            if (live) {
                for (int idx = within; idx < n; idx += L.slot_threads) {
                    int i = idx / d, v = idx % d;
                    float acc = 0.f, acc2 = 0.f;
                    for (int j = 0; j < d; ++j) {
                        float x = (float)X[i * d + j];
                        acc = fmaf(x, sA[v * d + j], acc);
                        if (dft) acc2 = fmaf(x, sB[v * d + j], acc2);
                    }
                    T[idx] = acc;
                    if (dft) T2[idx] = acc2;
                }
            }
            __syncthreads();
            // A5 second pass + A7 quantise + A8 zigzag
            if (live) {
                for (int idx = within; idx < n; idx += L.slot_threads) {
                    int u = idx / d, v = idx % d;
                    float acc = 0.f;
                    for (int i = 0; i < d; ++i) {
                        acc = fmaf(sA[u * d + i], T[i * d + v], acc);
                        if (dft) acc = fmaf(-sB[u * d + i], T2[i * d + v], acc);
                    }
                    float val = acc * a.t.qmult[idx];
                    float r = rintf(val);
                    float tol = a.t.qtol[idx];
                    if (!(g.flags & JB_FLAG_NO_REFINE) &&
                        fabsf(fabsf(val - r) - 0.5f) < tol + 2.4e-7f * fabsf(val)) {
                        r = (float)rint(jb_refine_coefficient(X, u, v, g, a.t));
                    }
                    int q = (int)r;
                    int zp = a.t.zz[idx];
                    if (MODE == 1) {
                        int qs = q > 32767 ? 32767 : (q < -32767 ? -32767 : q);
                        a.coeffs_out[((size_t)plane * g.nblocks + blk) * n + zp] = (int16_t)qs;
                    } else {
                        if (q > JB_MAX_AMP || q < -JB_MAX_AMP) {
                            int k = atomicAdd(&s_nbig, 1);
                            if (k < JB_BIGAMP_CAP) { s_big[k].blk = gi; s_big[k].pos = zp; s_big[k].amp = q; }
                            q = q > 0 ? 32767 : -32767;
                        }
                        ((int16_t*)(sCoef + gi * L.coefW))[zp] = (int16_t)q;
                    }
                }
            }
            __syncthreads();
        }
        if (MODE == 1) return;
    }

    // A9 + A10: one thread per block
    if (tid < JB_CHUNK) {
        unsigned len = 0;
        if (tid < nvalid) {
            int bad_pos, bad_run;
            const unsigned long long gblk = (unsigned long long)plane * g.nblocks + blk0 + tid;
            if (MODE == 2) {
                const int32_t* c = a.coeffs_in + gblk * n;
                len = jb_pack_block<int32_t>(c, n, sStage + tid * L.stageW, &bad_pos, &bad_run);
                if (bad_pos >= 0) jb_report_bad_code(jb_ctrl_status(a, P), gblk, bad_pos, bad_run, c[bad_pos]);
            } else {
                const int16_t* c = (const int16_t*)(sCoef + tid * L.coefW);
                len = jb_pack_block<int16_t>(c, n, sStage + tid * L.stageW, &bad_pos, &bad_run);
                if (bad_pos >= 0) {
                    long long amp = c[bad_pos];
                    int nb = jb_min(s_nbig, JB_BIGAMP_CAP);
                    for (int k = 0; k < nb; ++k)
                        if (s_big[k].blk == tid && s_big[k].pos == bad_pos) amp = s_big[k].amp;
                    jb_report_bad_code(jb_ctrl_status(a, P), gblk, bad_pos, bad_run, amp);
                }
            }
        }
        // warp scan of block lengths, then the chunk's place in the output
        unsigned incl = len;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
            if (tid >= o) incl += y;
        }
        const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
        s_blen[tid] = len;
        s_boff[tid] = incl - len;
        if (tid == 0) jb_record_chunk_len(a, P, chunk, total);
    }
    __syncthreads();
    // compacted copy of the chunk into its slot of the temporary buffer; the scan + gather
    // kernels below move it to its final place in the stream
    uint8_t* slot = jb_chunk_slot(a, chunk, s_boff[nvalid - 1] + s_blen[nvalid - 1]);
    for (int gi = 0; gi < nvalid; ++gi) {
        const uint8_t* sb = (const uint8_t*)(sStage + gi * L.stageW);
        uint8_t* dst = slot + s_boff[gi];
        for (unsigned j = tid; j < s_blen[gi]; j += JB_GENERIC_THREADS) dst[j] = sb[j];
    }
}

// ----------------------------------------------------------------------------------------------
// gather: every chunk from its slot to its place in the output stream.  No scan pass: a CTA (8 chunks, one per
// warp) derives the offset of its first chunk from the two levels of segment totals the transform kernel
// accumulated and the few chunk lengths in between.  The CTA that finishes last publishes the status words
// and leaves the control block and the totals zeroed for the next call on this workspace.
// ----------------------------------------------------------------------------------------------
#define JB_GATHER_WARPS 8

// The call's last kernel: hand the status words of set P to the caller, clean set P ^ 1 for the next call, flip
// the parity.  `slice` / `n_slices`: which part of the other set's segment totals this CTA zeroes.
__device__ __forceinline__ void jb_ctrl_finish(const JbFwdArgs& a, int P, bool leader, unsigned slice, unsigned n_slices,
                                               unsigned long long extra_error) {
    const int Q = P ^ 1;
    for (unsigned i = slice; i < a.n_seg1; i += n_slices) a.seg1[(size_t)Q * a.n_seg1 + i] = 0ull;
    for (unsigned i = slice; i < a.n_seg2; i += n_slices) a.seg2[(size_t)Q * a.n_seg2 + i] = 0ull;
    if (leader) {
        const unsigned long long* st = jb_ctrl_status(a, P);
        unsigned long long* sq = jb_ctrl_status(a, Q);
        const unsigned long long e0 = st[0];
        a.status_out[0] = e0 > extra_error ? e0 : extra_error;
        #pragma unroll
        for (int k = 1; k < JB_STATUS_WORDS; ++k) a.status_out[k] = st[k];
        #pragma unroll
        for (int k = 0; k < JB_STATUS_WORDS; ++k) sq[k] = k == 1 ? ~0ull : 0ull;
        *jb_ctrl_ticket(a, Q) = 0u;
        a.ctrl[0] = (unsigned)Q;
    }
}

__global__ void __launch_bounds__(JB_GATHER_WARPS * 32) jb_gather_chunks_kernel(JbFwdArgs a) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned c = blockIdx.x * JB_GATHER_WARPS + warp;
    const int P = (int)(__ldcg(a.ctrl + 1) & 1u);
    if (c < a.n_chunks) {
        const uint8_t* src = a.tmp_small + (size_t)c * JB_SLOT_STRIDE;   // 16-byte aligned, 1 KB + 32 B per slot
        const uint4* s16 = (const uint4*)src;
        const uint32_t* s32 = (const uint32_t*)src;
        // the first kilobyte of the slot is fetched before its length is known: the round trips (metadata, data)
        // overlap; an average chunk is ~0.6 KB
        uint4 pv[2];
        uint32_t pn[2];
        #pragma unroll
        for (int k = 0; k < 2; ++k) {
            pv[k] = __ldg(s16 + lane + 32 * k);
            pn[k] = __ldg(s32 + 4 * (lane + 32 * k) + 4);
        }
        // ---- offset of the chunk: second-level totals before its second-level segment, first-level totals before
        //      its first-level segment inside that, chunk lengths before it inside that: at most
        //      n_chunks / JB_SEG2 + JB_SEG2 / JB_SEG1 + JB_SEG1 values, a few per lane, all of them hot in L2 ----
        const unsigned long long* seg1 = a.seg1 + (size_t)P * a.n_seg1;
        const unsigned long long* seg2 = a.seg2 + (size_t)P * a.n_seg2;
        const unsigned g2 = c / JB_SEG2, g1 = c / JB_SEG1;
        unsigned long long base = 0ull;
        for (unsigned i = lane; i < g2; i += 32) base += __ldcg(seg2 + i);
        for (unsigned i = g2 * (JB_SEG2 / JB_SEG1) + lane; i < g1; i += 32) base += __ldcg(seg1 + i);
        for (unsigned i = g1 * JB_SEG1 + lane; i < c; i += 32) base += a.chunk_len[i];
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) base += __shfl_xor_sync(0xffffffffu, base, o);
        const unsigned len = a.chunk_len[c];
        if (len > JB_SLOT_SMALL) {                             // a dense chunk: it went to the worst-case-sized slot
            src = a.tmp + (size_t)c * a.chunk_cap;
            s16 = (const uint4*)src;
            s32 = (const uint32_t*)src;
            #pragma unroll
            for (int k = 0; k < 2; ++k) {
                pv[k] = __ldg(s16 + lane + 32 * k);
                pn[k] = __ldg(s32 + 4 * (lane + 32 * k) + 4);
            }
        }
        if (lane == 0 && c % (unsigned)a.g.cpp == 0) a.plane_off[c / (unsigned)a.g.cpp] = base;
        if (base + len <= a.out_cap) {                         // (otherwise flagged by CTA 0 below)
            uint8_t* dst = a.out + base;
            // head bytes up to a 4-byte boundary of dst; then every lane moves 16 bytes per pass: one
            // 128-bit load + the following word, funnel-shifted into four aligned words; then the tail
            const unsigned head = (unsigned)jb_min((int)len, (int)((4u - (unsigned)((uintptr_t)dst & 3u)) & 3u));
            if (lane < (int)head) dst[lane] = src[lane];
            const unsigned nwords = (len - head) >> 2;
            uint32_t* d32 = (uint32_t*)(dst + head);
            const unsigned sh = head * 8u;
            unsigned pass = 0;
            for (unsigned j = lane; j * 4u < nwords; j += 32, ++pass) {
                uint4 v;
                uint32_t nx;
                if (pass == 0) { v = pv[0]; nx = pn[0]; }
                else if (pass == 1) { v = pv[1]; nx = pn[1]; }
                else { v = __ldg(s16 + j); nx = __ldg(s32 + 4u * j + 4u); }
                uint32_t w0 = v.x, w1 = v.y, w2 = v.z, w3 = v.w;
                if (sh) {
                    w0 = __funnelshift_r(v.x, v.y, sh); w1 = __funnelshift_r(v.y, v.z, sh);
                    w2 = __funnelshift_r(v.z, v.w, sh); w3 = __funnelshift_r(v.w, nx, sh);
                }
                const unsigned w = 4u * j;
                d32[w] = w0;
                if (w + 1 < nwords) d32[w + 1] = w1;
                if (w + 2 < nwords) d32[w + 2] = w2;
                if (w + 3 < nwords) d32[w + 3] = w3;
            }
            const unsigned tail0 = head + nwords * 4u;
            if (tail0 + lane < len) dst[tail0 + lane] = src[tail0 + lane];
        }
    }
    // ---- end of the call: grand total and capacity check (CTA 0), status out, the other set cleaned ----
    if (warp == 0) {
        unsigned long long err = 0ull;
        if (blockIdx.x == 0) {
            const unsigned long long* seg2 = a.seg2 + (size_t)P * a.n_seg2;
            unsigned long long tot = 0ull;
            for (unsigned i = lane; i < a.n_seg2; i += 32) tot += __ldcg(seg2 + i);
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            if (lane == 0) a.plane_off[a.n_planes] = tot;
            if (tot > a.out_cap) err = (unsigned long long)(-JB_ERR_OUT_CAPACITY);
        }
        if (lane == 0) jb_ctrl_finish(a, P, blockIdx.x == 0, blockIdx.x, gridDim.x, err);
    }
}

// calls that end without a gather pass (stage entry point: coefficients only)
__global__ void __launch_bounds__(256) jb_finish_kernel(JbFwdArgs a) {
    const int P = (int)(__ldcg(a.ctrl + 1) & 1u);
    jb_ctrl_finish(a, P, threadIdx.x == 0, threadIdx.x, blockDim.x, 0ull);
}

__global__ void __launch_bounds__(256) jb_init_ctrl_kernel(unsigned* ctrl, unsigned long long* seg, size_t n_seg_words) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x < JB_CTRL_BYTES / 4) {
        unsigned v = 0u;
        // status[P][1] = all ones ("no bad code yet") in both sets
        const unsigned w = threadIdx.x - JB_CTRL_STATUS_OFF / 4;
        if (threadIdx.x >= JB_CTRL_STATUS_OFF / 4 && w < 4u * JB_STATUS_WORDS && (w % (2u * JB_STATUS_WORDS)) / 2u == 1u) v = 0xFFFFFFFFu;
        ctrl[threadIdx.x] = v;
    }
    for (size_t i = t; i < n_seg_words; i += (size_t)gridDim.x * blockDim.x) seg[i] = 0ull;
}

cudaError_t jb_launch_init_ctrl(unsigned* ctrl, unsigned long long* seg, size_t n_seg_words, cudaStream_t s) {
    size_t blocks = (n_seg_words + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 1024) blocks = 1024;
    jb_init_ctrl_kernel<<<(unsigned)blocks, 256, 0, s>>>(ctrl, seg, n_seg_words);
    return cudaGetLastError();
}

cudaError_t jb_launch_gather(const JbFwdArgs& a, cudaStream_t s) {
    jb_prof_mark(1, s);                                // the fused forward kernel was launched just before
    if (a.n_chunks == 0) return jb_launch_finish(a, s);
    jb_gather_chunks_kernel<<<(a.n_chunks + JB_GATHER_WARPS - 1) / JB_GATHER_WARPS, JB_GATHER_WARPS * 32, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t jb_launch_finish(const JbFwdArgs& a, cudaStream_t s) {
    jb_finish_kernel<<<1, 256, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t jb_launch_fwd_generic(const JbFwdArgs& a, int mode, cudaStream_t s) {
    const bool dft = a.g.transform == JB_TRANSFORM_DFT;
    size_t smem = jb_fwd_generic_smem_bytes(a.g.d, dft);
    cudaError_t e;
    if (a.n_chunks == 0) return cudaSuccess;
    switch (mode) {
    case 0:
        e = cudaFuncSetAttribute(jb_fwd_generic_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        jb_fwd_generic_kernel<0><<<a.n_chunks, JB_GENERIC_THREADS, smem, s>>>(a);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        return jb_launch_gather(a, s);
    case 1:
        e = cudaFuncSetAttribute(jb_fwd_generic_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        jb_fwd_generic_kernel<1><<<a.n_chunks, JB_GENERIC_THREADS, smem, s>>>(a);
        break;
    default:
        e = cudaFuncSetAttribute(jb_fwd_generic_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        jb_fwd_generic_kernel<2><<<a.n_chunks, JB_GENERIC_THREADS, smem, s>>>(a);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        return jb_launch_gather(a, s);
    }
    return cudaGetLastError();
}
