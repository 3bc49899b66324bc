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
    return jb_refine_f64(X, u, v, g.d, g.bs, g.transform, g.qmode, t.fA64, t.fB64, t.qrecip[u * g.d + v]);
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
    const int blk0 = (chunk % g.cpp) * g.chunk;
    const int nvalid = jb_min(g.chunk, g.nblocks - blk0);

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
// device-wide exclusive scan of the chunk lengths in ONE launch, then gather
//   scan:   one CTA per segment of JB_SCAN_SEG chunks: offsets inside the segment + segment total; the CTA
//           that finishes last scans the segment totals in place (exclusive bases), writes the grand total and
//           checks the capacity
//   gather: one warp per chunk: copy slot -> out[base], record the start of every plane; CTA 0 ends the call:
//           status words out, the other control set cleaned, parity flipped
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(JB_SCAN_SEG) jb_scan_kernel(JbFwdArgs a, unsigned n_seg) {
    __shared__ unsigned s_warp[33];
    __shared__ unsigned long long s_carry;
    __shared__ int s_last;
    jb_pdl_trigger();
    jb_pdl_wait();
    const unsigned c = blockIdx.x * JB_SCAN_SEG + threadIdx.x;
    const unsigned len = c < a.n_chunks ? a.chunk_len[c] : 0u;
    unsigned total;
    const unsigned ex = jb_block_excl_scan(len, s_warp, &total);
    if (c < a.n_chunks) a.chunk_off[c] = ex;
    if (threadIdx.x == 0) {
        a.seg_total[blockIdx.x] = total;
        __threadfence();
        s_last = atomicAdd(a.ctrl + 4, 1u) == gridDim.x - 1u ? 1 : 0;
        s_carry = 0ull;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (unsigned base = 0; base < n_seg; base += JB_SCAN_SEG) {
        const unsigned i = base + threadIdx.x;
        // segment totals fit 32 bits but their running sum does not: scan the low 20 bits and the rest separately
        const unsigned long long v = i < n_seg ? __ldcg(a.seg_total + i) : 0ull;
        unsigned tlo, thi;
        const unsigned lo = jb_block_excl_scan((unsigned)(v & 0xFFFFFu), s_warp, &tlo);
        const unsigned hi = jb_block_excl_scan((unsigned)(v >> 20), s_warp, &thi);
        const unsigned long long carry = s_carry;
        if (i < n_seg) a.seg_total[i] = carry + (((unsigned long long)hi) << 20) + lo;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + (((unsigned long long)thi) << 20) + tlo;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int P = (int)(__ldcg(a.ctrl + 1) & 1u);
        a.plane_off[a.n_planes] = s_carry;
        if (s_carry > a.out_cap) jb_set_error(jb_ctrl_status(a, P), JB_ERR_OUT_CAPACITY);
        a.ctrl[4] = 0u;
    }
}

#define JB_GATHER_WARPS 8

// The call's last kernel: hand the status words of set P to the caller, clean set P ^ 1 for the next call, flip
// the parity (one thread).
__device__ __forceinline__ void jb_ctrl_finish(const JbFwdArgs& a, int P) {
    const int Q = P ^ 1;
    const unsigned long long* st = jb_ctrl_status(a, P);
    unsigned long long* sq = jb_ctrl_status(a, Q);
    #pragma unroll
    for (int k = 0; k < JB_STATUS_WORDS; ++k) a.status_out[k] = __ldcg(st + k);
    #pragma unroll
    for (int k = 0; k < JB_STATUS_WORDS; ++k) sq[k] = k == 1 ? ~0ull : 0ull;
    *jb_ctrl_ticket(a, Q) = 0u;
    a.ctrl[0] = (unsigned)Q;
}

// one warp: chunk c from its slot to byte `base` of the output
__device__ __forceinline__ void jb_gather_chunk(const JbFwdArgs& a, unsigned c, int lane, unsigned long long base, bool base_known,
                                                const unsigned* s_off) {
    // 16-byte aligned slots: the dense array of 1 KB + 32 B ones, or the worst-case sized ones
    const uint8_t* src = a.slots_always_big ? a.tmp + (size_t)c * a.chunk_cap : a.tmp_small + (size_t)c * JB_SLOT_STRIDE;
    const uint4* s16 = (const uint4*)src;
    const uint32_t* s32 = (const uint32_t*)src;
    // the first kilobyte of the slot is fetched before its length is known: the two round trips
    // (metadata, data) overlap; an average chunk is ~0.6 KB
    uint4 pv[2];
    uint32_t pn[2];
    #pragma unroll
    for (int k = 0; k < 2; ++k) {
        pv[k] = __ldg(s16 + lane + 32 * k);
        pn[k] = __ldg(s32 + 4 * (lane + 32 * k) + 4);
    }
    const unsigned len = a.chunk_len[c];
    if (len > JB_SLOT_SMALL && !a.slots_always_big) {          // a dense chunk: it went to the worst-case-sized slot
        src = a.tmp + (size_t)c * a.chunk_cap;
        s16 = (const uint4*)src;
        s32 = (const uint32_t*)src;
        #pragma unroll
        for (int k = 0; k < 2; ++k) {
            pv[k] = __ldg(s16 + lane + 32 * k);
            pn[k] = __ldg(s32 + 4 * (lane + 32 * k) + 4);
        }
    }
    if (!base_known) base = s_off ? (unsigned long long)s_off[c] : a.seg_total[c / JB_SCAN_SEG] + a.chunk_off[c];
    if (lane == 0 && c % (unsigned)a.g.cpp == 0) a.plane_off[c / (unsigned)a.g.cpp] = base;
    if (base + len > a.out_cap) return;                       // flagged by the scan
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

__global__ void __launch_bounds__(JB_GATHER_WARPS * 32) jb_gather_chunks_kernel(JbFwdArgs a) {
    const int lane = threadIdx.x & 31;
    const unsigned c = blockIdx.x * JB_GATHER_WARPS + (threadIdx.x >> 5);
    jb_pdl_trigger();
    jb_pdl_wait();
    if (blockIdx.x == 0 && threadIdx.x == 0) jb_ctrl_finish(a, (int)(__ldcg(a.ctrl + 1) & 1u));
    if (c >= a.n_chunks) return;
    jb_gather_chunk(a, c, lane, 0ull, false, nullptr);
}

// Scan and gather in one launch for calls of a few frames (up to JB_TAIL_SMALL chunks), where a launch costs as much
// as the work: every CTA scans all the chunk lengths for itself (a few kilobytes out of L2) and gathers its own eight
// chunks; CTA 0 also ends the call.
#define JB_TAIL_SMALL 2048u
__global__ void __launch_bounds__(JB_GATHER_WARPS * 32) jb_scan_gather_small_kernel(JbFwdArgs a) {
    __shared__ unsigned s_off[JB_TAIL_SMALL];
    __shared__ unsigned s_warp[33];
    const int lane = threadIdx.x & 31;
    jb_pdl_trigger();
    jb_pdl_wait();
    constexpr unsigned PER = JB_TAIL_SMALL / (JB_GATHER_WARPS * 32);
    unsigned len[PER], sum = 0;
    #pragma unroll
    for (unsigned k = 0; k < PER; ++k) {
        const unsigned c = threadIdx.x * PER + k;
        len[k] = c < a.n_chunks ? __ldcg(a.chunk_len + c) : 0u;
        sum += len[k];
    }
    unsigned total;
    unsigned ex = jb_block_excl_scan(sum, s_warp, &total);
    #pragma unroll
    for (unsigned k = 0; k < PER; ++k) { s_off[threadIdx.x * PER + k] = ex; ex += len[k]; }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int P = (int)(__ldcg(a.ctrl + 1) & 1u);
        a.plane_off[a.n_planes] = total;
        if (total > a.out_cap) jb_set_error(jb_ctrl_status(a, P), JB_ERR_OUT_CAPACITY);
        jb_ctrl_finish(a, P);
    }
    const unsigned c = blockIdx.x * JB_GATHER_WARPS + (threadIdx.x >> 5);
    if (c >= a.n_chunks) return;
    jb_gather_chunk(a, c, lane, 0ull, false, s_off);
}

// calls that end without a gather pass (stage entry point: coefficients only)
__global__ void jb_finish_kernel(JbFwdArgs a) { jb_ctrl_finish(a, (int)(__ldcg(a.ctrl + 1) & 1u)); }

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
    JB_LAUNCH((jb_init_ctrl_kernel), (unsigned)blocks, 256, 0, s, ctrl, seg, n_seg_words);
    return cudaGetLastError();
}

cudaError_t jb_launch_gather(const JbFwdArgs& a, cudaStream_t s) {
    jb_prof_mark(1, s);                                // the fused forward kernel was launched just before
    if (a.n_chunks == 0) return jb_launch_finish(a, s);
    const bool pdl = (a.g.flags & JB_FLAG_PDL) != 0;
    if (a.n_chunks <= JB_TAIL_SMALL)
        return jb_launch_ex(jb_scan_gather_small_kernel, dim3((a.n_chunks + JB_GATHER_WARPS - 1) / JB_GATHER_WARPS),
                            dim3(JB_GATHER_WARPS * 32), 0, s, pdl, a);
    const unsigned n_seg = (a.n_chunks + JB_SCAN_SEG - 1) / JB_SCAN_SEG;
    cudaError_t e = jb_launch_ex(jb_scan_kernel, dim3(n_seg), dim3(JB_SCAN_SEG), 0, s, pdl, a, n_seg);
    if (e != cudaSuccess) return e;
    return jb_launch_ex(jb_gather_chunks_kernel, dim3((a.n_chunks + JB_GATHER_WARPS - 1) / JB_GATHER_WARPS),
                        dim3(JB_GATHER_WARPS * 32), 0, s, pdl, a);
}

cudaError_t jb_launch_finish(const JbFwdArgs& a, cudaStream_t s) {
    JB_LAUNCH((jb_finish_kernel), 1, 1, 0, s, a);
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
        JB_LAUNCH((jb_fwd_generic_kernel<0>), a.n_chunks, JB_GENERIC_THREADS, smem, s, a);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        return jb_launch_gather(a, s);
    case 1:
        e = cudaFuncSetAttribute(jb_fwd_generic_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        JB_LAUNCH((jb_fwd_generic_kernel<1>), a.n_chunks, JB_GENERIC_THREADS, smem, s, a);
        break;
    default:
        e = cudaFuncSetAttribute(jb_fwd_generic_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        JB_LAUNCH((jb_fwd_generic_kernel<2>), a.n_chunks, JB_GENERIC_THREADS, smem, s, a);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        return jb_launch_gather(a, s);
    }
    return cudaGetLastError();
}
