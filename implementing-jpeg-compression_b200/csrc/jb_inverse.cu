// jb_inverse.cu -- decompress direction, generic kernel.
//
//   jb_inv_generic_kernel : any (block_size, dct_size <= 32, DCT | DFT, quantiser); one CTA per
//                           chunk of 32 blocks.  Reference stages 8-0 inverted and fused:
//                           rle_byte_stream.py:60-88, run_length_encoding.py:64-78,31-41,
//                           zigzag_order.py:101-119, quantization.py:20-30 (quantizers.py restore),
//                           basis_change.py:28-43 (transforms.py:40-44,60-69, np.round -> int),
//                           normalization.py:10-14 (clamp), dct_padding.py:11-21 (crop),
//                           subsampling.py:13-14 (util.inflate), padding.py:14-16 (crop),
//                           and the uint8 cast of pipeline/__init__.py:120-122.
//   Block offsets come from jb_framing.cu.
#include "jb_common.cuh"
#include "jb_inverse.cuh"

struct JbInvSmemLayout {
    int slot_threads, nsub;
    int coefW;
    size_t off_A, off_B, off_dq, off_Y, off_P, off_P2, off_coef, total;
};

__host__ __device__ inline JbInvSmemLayout jb_inv_smem_layout(int d, bool dft) {
    JbInvSmemLayout L;
    int n = d * d;
    int st = (n + 31) / 32 * 32;
    if (st > JB_INV_GENERIC_THREADS) st = JB_INV_GENERIC_THREADS;
    L.slot_threads = st;
    L.nsub = JB_INV_GENERIC_THREADS / st;
    L.coefW = ((n + 1) / 2) | 1;
    size_t o = 0;
    L.off_A = o;    o += (size_t)n * 4;
    L.off_B = o;    o += dft ? (size_t)n * 4 : 0;
    L.off_dq = o;   o += (size_t)n * 4;
    L.off_Y = o;    o += (size_t)L.nsub * n * 4;
    L.off_P = o;    o += (size_t)L.nsub * n * 4;
    L.off_P2 = o;   o += dft ? (size_t)L.nsub * n * 4 : 0;
    L.off_coef = o; o += (size_t)JB_CHUNK * L.coefW * 4;
    L.total = o;
    return L;
}

size_t jb_inv_generic_smem_bytes(int d, bool dft) { return jb_inv_smem_layout(d, dft).total; }

// end of a decompress call whose last transform kernel has no epilogue of its own (everything but the fused 8x8 kernel)
__global__ void jb_dec_finish_kernel(JbInvArgs a) { jb_dec_publish_and_clean(a); }

cudaError_t jb_launch_dec_finish(const JbInvArgs& a, cudaStream_t s) {
    JB_LAUNCH((jb_dec_finish_kernel), 1, 1, 0, s, a);
    return cudaGetLastError();
}

template <int MODE>
__global__ void __launch_bounds__(JB_INV_GENERIC_THREADS)
jb_inv_generic_kernel(const JbInvArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const JbGeom& g = a.g;
    const int d = g.d, n = g.n;
    const bool dft = g.transform == JB_TRANSFORM_DFT;
    const JbInvSmemLayout L = jb_inv_smem_layout(d, dft);
    float* sA = (float*)(smem + L.off_A);
    float* sB = (float*)(smem + L.off_B);
    float* sDq = (float*)(smem + L.off_dq);
    float* sY = (float*)(smem + L.off_Y);
    float* sP = (float*)(smem + L.off_P);
    float* sP2 = (float*)(smem + L.off_P2);
    uint32_t* sCoef = (uint32_t*)(smem + L.off_coef);

    const int tid = threadIdx.x;
    const unsigned chunk = blockIdx.x;
    if (chunk >= a.n_chunks) return;
    const int plane = chunk / g.cpp;
    const int blk0 = (chunk % g.cpp) * g.chunk;
    const int nvalid = jb_min(g.chunk, g.nblocks - blk0);

    if (MODE != 1) {
        for (int i = tid; i < n; i += JB_INV_GENERIC_THREADS) {
            sA[i] = a.t.iA[i];
            if (dft) sB[i] = a.t.iB[i];
            sDq[i] = a.t.dqmult[i];
        }
    }
    for (int i = tid; i < JB_CHUNK * L.coefW; i += JB_INV_GENERIC_THREADS) sCoef[i] = 0u;
    __syncthreads();

    if (MODE != 2) {
        // A10/A9 inverted: one thread per block walks its codes
        if (tid < nvalid) {
            const unsigned long long len = a.plane_len[plane];
            const uint8_t* stream = a.in + a.plane_off[plane];
            const unsigned start = a.block_start[(size_t)plane * g.nblocks + blk0 + tid];
            const unsigned limit = len > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned)len;
            int16_t* row = (int16_t*)(sCoef + tid * L.coefW);
            // (a stream must lie inside the caller's buffer: device-supplied offsets are not trusted)
            const bool inside = a.plane_off[plane] <= a.in_bytes && len <= a.in_bytes - a.plane_off[plane];
            int rc = (inside && start < limit)
                ? jb_decode_block<int16_t, uint16_t>(stream, start, limit, n, row,
                                                     MODE == 1 ? (const uint16_t*)nullptr : a.t.izz)
                : JB_PARSE_BAD;
            if (rc != JB_PARSE_OK) jb_set_error(a.status, JB_ERR_BAD_STREAM);
        }
    } else {
        for (int idx = tid; idx < nvalid * n; idx += JB_INV_GENERIC_THREADS) {
            int gi = idx / n, zp = idx % n;
            int16_t c = a.coeffs_in[((size_t)plane * g.nblocks + blk0 + gi) * n + zp];
            ((int16_t*)(sCoef + gi * L.coefW))[a.t.izz[zp]] = c;
        }
    }
    __syncthreads();

    if (MODE == 1) {
        for (int idx = tid; idx < nvalid * n; idx += JB_INV_GENERIC_THREADS) {
            int gi = idx / n, zp = idx % n;
            a.coeffs_out[((size_t)plane * g.nblocks + blk0 + gi) * n + zp] =
                ((const int16_t*)(sCoef + gi * L.coefW))[zp];
        }
        return;
    }

    uint8_t* dst = a.planes_out + (size_t)plane * a.plane_stride;
    const int slot = tid / L.slot_threads, within = tid % L.slot_threads;
    for (int base = 0; base < nvalid; base += L.nsub) {
        const int gi = base + slot;
        const bool live = slot < L.nsub && gi < nvalid;
        float* Y = sY + slot * n;
        float* P = sP + slot * n;
        float* P2 = sP2 + slot * n;
        const int blk = blk0 + gi;
        const int by = blk / g.hb, bx = blk % g.hb;
        // A7 inverted: coefficient * quantiser step (exact integers in fp32)
        if (live) {
            const int16_t* row = (const int16_t*)(sCoef + gi * L.coefW);
            for (int idx = within; idx < n; idx += L.slot_threads) Y[idx] = (float)row[idx] * sDq[idx];
        }
        __syncthreads();
        // A6 first pass along v: P[u][c] = sum_v Y[u][v] iA[c][v]
        if (live) {
            for (int idx = within; idx < n; idx += L.slot_threads) {
                int u = idx / d, c = idx % d;
                float acc = 0.f, acc2 = 0.f;
                for (int v = 0; v < d; ++v) {
                    float y = Y[u * d + v];
                    acc = fmaf(y, sA[c * d + v], acc);
                    if (dft) acc2 = fmaf(y, sB[c * d + v], acc2);
                }
                P[idx] = acc;
                if (dft) P2[idx] = acc2;
            }
        }
        __syncthreads();
        // second pass along u, round, clamp, crop, replicate
        if (live) {
            for (int idx = within; idx < n; idx += L.slot_threads) {
                int r = idx / d, c = idx % d;
                float acc = 0.f;
                for (int u = 0; u < d; ++u) {
                    acc = fmaf(sA[r * d + u], P[u * d + c], acc);
                    if (dft) acc = fmaf(-sB[r * d + u], P2[u * d + c], acc);
                }
                float x = rintf(acc);                                   // basis_change.py:43
                x = fminf(fmaxf(x, 0.f), 255.f);                        // normalization.py:10-14
                const int si = by * d + r, sj = bx * d + c;
                if (si < g.H1 && sj < g.W1) {                           // dct_padding.py:11-21
                    const uint8_t pix = (uint8_t)x;
                    for (int di = 0; di < g.bs; ++di) {
                        int yy = si * g.bs + di;
                        if (yy >= g.H) break;                           // padding.py:14-16
                        uint8_t* orow = dst + (size_t)yy * a.row_pitch;
                        for (int dj = 0; dj < g.bs; ++dj) {
                            int xx = sj * g.bs + dj;
                            if (xx >= g.W) break;
                            orow[xx] = pix;
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
}

cudaError_t jb_launch_inv_generic(const JbInvArgs& a, int mode, cudaStream_t s) {
    const bool dft = a.g.transform == JB_TRANSFORM_DFT;
    size_t smem = jb_inv_generic_smem_bytes(a.g.d, dft);
    cudaError_t e;
    if (a.n_chunks == 0) return cudaSuccess;
    switch (mode) {
    case 0:
        e = cudaFuncSetAttribute(jb_inv_generic_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        JB_LAUNCH((jb_inv_generic_kernel<0>), a.n_chunks, JB_INV_GENERIC_THREADS, smem, s, a);
        break;
    case 1:
        e = cudaFuncSetAttribute(jb_inv_generic_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        JB_LAUNCH((jb_inv_generic_kernel<1>), a.n_chunks, JB_INV_GENERIC_THREADS, smem, s, a);
        break;
    default:
        e = cudaFuncSetAttribute(jb_inv_generic_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        JB_LAUNCH((jb_inv_generic_kernel<2>), a.n_chunks, JB_INV_GENERIC_THREADS, smem, s, a);
        break;
    }
    return cudaGetLastError();
}
