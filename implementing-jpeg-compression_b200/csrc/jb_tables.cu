// jb_tables.cu -- parameter validation, geometry and the per-call table builder.
#include <math.h>
#include <mutex>

#include "jb_common.cuh"

// quantizers.py:35-42 (the standard JPEG luminance table)
__constant__ unsigned char jb_qtable_c[64] = {
    16, 11, 10, 16, 24, 40, 51, 61,
    12, 12, 14, 19, 26, 58, 60, 55,
    14, 13, 16, 24, 40, 57, 69, 56,
    14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77,
    24, 35, 55, 64, 81, 104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101,
    72, 92, 95, 98, 112, 100, 103, 99};

static inline int jb_ceil_div(int a, int b) { return (a + b - 1) / b; }

int jb_make_geom(const jb_params* p, JbGeom* g) {
    if (!p || !g) return JB_ERR_BAD_PARAM;
    if (p->height < 0 || p->width < 0) return JB_ERR_BAD_PARAM;
    if (p->height == 0 || p->width == 0) return JB_ERR_EMPTY_ARRAY;        // util.py:30-31
    if (p->block_size < 1 || p->dct_size < 1) return JB_ERR_BAD_PARAM;
    if (p->transform != JB_TRANSFORM_DCT && p->transform != JB_TRANSFORM_DFT) return JB_ERR_BAD_PARAM;
    if (p->qmode < JB_Q_NONE || p->qmode > JB_Q_QTABLE) return JB_ERR_BAD_QUANTIZATION;
    if (p->qmode == JB_Q_QTABLE && p->dct_size != 8) return JB_ERR_BAD_QUANTIZATION;   // pipeline/__init__.py:62-63
    if (p->qmode == JB_Q_DIVIDE && p->qparam == 0) return JB_ERR_BAD_QUANTIZATION;
    if (p->qmode == JB_Q_DISCARD && p->qparam < 0) return JB_ERR_BAD_QUANTIZATION;
    if (p->dct_size > JB_MAX_DCT_SIZE || p->block_size > JB_MAX_BLOCK_SIZE) return JB_ERR_UNSUPPORTED;
    g->H = p->height; g->W = p->width; g->bs = p->block_size; g->d = p->dct_size;
    g->n = g->d * g->d;
    g->H1 = jb_ceil_div(g->H, g->bs); g->W1 = jb_ceil_div(g->W, g->bs);
    g->vb = jb_ceil_div(g->H1, g->d); g->hb = jb_ceil_div(g->W1, g->d);
    g->H2 = g->vb * g->d; g->W2 = g->hb * g->d;
    long long nb = (long long)g->vb * g->hb;
    if (nb > 0x3FFFFFFF) return JB_ERR_UNSUPPORTED;
    g->nblocks = (int)nb;
    g->maxblk = jb_max_block_bytes(g->n);
    g->chunk = jb_chunk_blocks(g->d);
    g->cpp = jb_ceil_div(g->nblocks, g->chunk);
    g->transform = p->transform; g->qmode = p->qmode; g->qparam = p->qparam; g->flags = p->flags;
    return JB_OK;
}

// Zigzag position of (i, j) in an n x n block (pipeline/zigzag_order.py:27-43,55-79):
// anti-diagonals s = i + j in increasing order; even diagonals run from the bottom-left
// element up to the right, odd diagonals are reversed.
__device__ __forceinline__ int jb_zigzag_pos(int i, int j, int n) {
    int s = i + j;
    int before = (s < n) ? s * (s + 1) / 2 : n * n - (2 * n - 1 - s) * (2 * n - s) / 2;
    int i_lo = s < n ? 0 : s - n + 1, i_hi = s < n ? s : n - 1;
    // unreversed order starts at the largest i (bottom-left) and walks up-right
    int k = (s % 2 == 0) ? (i_hi - i) : (i - i_lo);
    return before + k;
}

// The un-normalised DCT-II matrix of transforms.py:4-11, evaluated on the HOST with libm's cos (the function
// numpy's np.cos agrees with to the last bit on the arguments the reference forms; CUDA's device cos may differ
// by an ulp) and handed to the builder kernel by value: the float64 re-evaluation path then works on the
// reference's own matrix.
struct JbHostMatrix { double a[JB_MAX_DCT_SIZE * JB_MAX_DCT_SIZE]; };

// One thread per (k, m) table entry; one block.  All arithmetic in double, written
// the way the reference writes it so that the float64 re-evaluation path reproduces
// its values (transforms.py:4-26, quantizers.py:27-28,47-53).
__global__ void jb_build_tables_kernel(JbGeom g, JbTables t, const __grid_constant__ JbHostMatrix hm) {
    const int d = g.d, n = g.n;
    __shared__ double rowabs_f[JB_MAX_DCT_SIZE];   // sum_m |A[k][m]| (+|B[k][m]| for the DFT)
    __shared__ double norm2[JB_MAX_DCT_SIZE];      // |c_k|^2
    for (int k = threadIdx.x; k < d; k += blockDim.x) {
        double sa = 0.0, s2 = 0.0;
        for (int m = 0; m < d; ++m) {
            double a, b = 0.0;
            if (g.transform == JB_TRANSFORM_DCT) {
                a = hm.a[k * d + m];
            } else {
                // (k*m mod d) keeps the angle small; cospi/sinpi make the trivial twiddles
                // (0, +-1) exact, as they are inside an FFT (np.fft.fft2, basis_change.py:23)
                int r = (k * m) % d;
                a = cospi(2.0 * r / d);
                b = sinpi(2.0 * r / d);
            }
            sa += fabs(a) + fabs(b);
            s2 += a * a;
        }
        rowabs_f[k] = sa;
        norm2[k] = s2;
    }
    __syncthreads();
    const double bs2 = (double)g.bs * (double)g.bs;
    const double M = 255.0 * bs2;                   // largest box sum
    const double u = 5.9604644775390625e-08;        // 2^-24
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        int k = idx / d, m = idx % d;
        double a, b = 0.0, ia, ib = 0.0;
        if (g.transform == JB_TRANSFORM_DCT) {
            a = hm.a[idx];
            // inverse: x = W (Dinv y), W[m][k] = C[k][m]/|c_k|, Dinv[k] = 1/|c_k| (transforms.py:14-26,40-44)
            double ct = hm.a[m * d + k];                      // C[m][k] read as entry (row k of B = sample k, col m = freq m)
            double nm = sqrt(norm2[m]);
            ia = (ct / nm) * (1.0 / nm);
        } else {
            int r = (k * m) % d;
            a = cospi(2.0 * r / d);
            b = sinpi(2.0 * r / d);
            ia = a / d;                                       // Re ifft2 = (Cc Q Cc - S Q S) / N^2
            ib = b / d;
        }
        t.fA64[idx] = a;  t.fA[idx] = (float)a;
        t.fB64[idx] = b;  t.fB[idx] = (float)b;
        t.iA[idx] = (float)ia;  t.iB[idx] = (float)ib;

        // quantiser for natural index idx = u*d + v  (k = u, m = v here)
        double recip, dq;
        bool dropped = false;
        if (g.qmode == JB_Q_QTABLE) {
            recip = 1.0 / (double)jb_qtable_c[idx];           // quantizers.py:47-49 multiplies by 1.0/q
            dq = (double)jb_qtable_c[idx];
        } else if (g.qmode == JB_Q_DIVIDE) {
            recip = (double)g.qparam;                         // quantizers.py:27-28 divides by float(divisor)
            dq = (double)g.qparam;
        } else {
            recip = 1.0;
            dq = 1.0;
            if (g.qmode == JB_Q_DISCARD && (k >= g.qparam || m >= g.qparam)) dropped = true;   // quantizers.py:16-20
        }
        double mult = (g.qmode == JB_Q_DIVIDE) ? 1.0 / (recip * bs2) : recip / bs2;
        if (dropped) { mult = 0.0; recip = 0.0; }
        t.qrecip[idx] = recip;
        t.qmult[idx] = (float)mult;
        t.dqmult[idx] = (float)dq;
        // fp32 error bound of the two-pass contraction on integer sums <= M:
        //   |err(Y[u][v])| <= 2 (d + 1) u R_u R_v M   (R = absolute row sum), doubled for the DFT's two products
        double tolY = 2.0 * (d + 1) * u * rowabs_f[k] * rowabs_f[m] * M;
        t.qtol[idx] = dropped ? -1.0f : (float)(tolY * fabs(mult) + 1e-7);
        int zp = jb_zigzag_pos(k, m, d);
        t.zz[idx] = (uint16_t)zp;
        t.izz[zp] = (uint16_t)idx;
    }
}

cudaError_t jb_launch_build_tables(const JbGeom& g, const JbTables& t, cudaStream_t s) {
    static JbHostMatrix hm;                        // (8 KB: not on the stack; the launch copies it, so reuse is safe)
    static int hm_d = 0;
    static std::mutex hm_lock;
    std::lock_guard<std::mutex> guard(hm_lock);
    if (g.transform == JB_TRANSFORM_DCT && hm_d != g.d) {
        const double PI = 3.141592653589793;       // np.pi
        for (int k = 0; k < g.d; ++k)
            for (int m = 0; m < g.d; ++m) hm.a[k * g.d + m] = cos(PI / g.d * (m + 0.5) * k);     // transforms.py:9-10
        hm_d = g.d;
    }
    JB_LAUNCH((jb_build_tables_kernel), 1, 256, 0, s, g, t, hm);
    return cudaGetLastError();
}
