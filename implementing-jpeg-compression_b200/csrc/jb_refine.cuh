// jb_refine.cuh -- float64 re-evaluation of one coefficient that the fp32 transform left within its proven
// error bound of a rounding tie, in the operation order of the reference so that the rounded integer is the
// reference's (np.round, quantizers.py:6,17,28,49).
//
//   DCT (transforms.py:46-58): rows first, M[i][v] = sum_j x[i][j] C[v][j], then Y[u][v] = sum_i M[i][v] C[u][i];
//       every sum term by term in ascending index with separate multiply and add -- the order oracle/ref_port.py
//       fixes for ndarray.dot (whose own order belongs to the BLAS build of the machine).
//   DFT, dct_size 8 (basis_change.py:20-25): np.fft.fft2 = pocketfft's complex radix-8 pass along the rows, then
//       along the columns.  jb_pf_pass8 repeats that pass operation by operation (pinned bitwise against numpy by
//       tests/test_pocketfft_model.py), so the real part it returns is numpy's to the last bit.
//   DFT, other sizes: direct sums with exact trivial twiddles (pocketfft's other radices are not restated).
// Every operation is an explicit round-to-nearest intrinsic: nvcc must not contract a*b+c into an FMA here.
#pragma once
#include "jb_common.cuh"

struct JbC64 { double r, i; };

// sum / (bs * bs) as the reference's np.mean rounds it
struct JbBoxMean {
    double k;
    bool pow2;
    __device__ __forceinline__ explicit JbBoxMean(int bs) {
        const int c = bs * bs;
        pow2 = (c & (c - 1)) == 0;
        k = pow2 ? 1.0 / (double)c : (double)c;
    }
    __device__ __forceinline__ double operator()(double sum) const { return pow2 ? __dmul_rn(sum, k) : __ddiv_rn(sum, k); }
};

__device__ __forceinline__ JbC64 jb_cadd(JbC64 a, JbC64 b) { return {__dadd_rn(a.r, b.r), __dadd_rn(a.i, b.i)}; }
__device__ __forceinline__ JbC64 jb_csub(JbC64 a, JbC64 b) { return {__dsub_rn(a.r, b.r), __dsub_rn(a.i, b.i)}; }
__device__ __forceinline__ JbC64 jb_rot90(JbC64 a) { return {a.i, -a.r}; }                       // forward: times -i

// pocketfft (C++, as built into numpy >= 2.0) cfftp::pass8<fwd = true> with ido = l1 = 1, the whole plan of a
// length-8 transform: PM / PMINPLACE / ROTX90 / ROTX45 / ROTX135 in its order, hsqt2 = sqrt(1/2) rounded.
__device__ __forceinline__ void jb_pf_pass8_body(const JbC64* c, JbC64* ch) {
    const double h = 0.70710678118654752440;
    JbC64 a0, a1, a2, a3, a4, a5, a6, a7, s, t;
    a1 = jb_cadd(c[1], c[5]); a5 = jb_csub(c[1], c[5]);
    a3 = jb_cadd(c[3], c[7]); a7 = jb_csub(c[3], c[7]);
    s = jb_cadd(a1, a3); a3 = jb_csub(a1, a3); a1 = s;
    a3 = jb_rot90(a3);
    a7 = jb_rot90(a7);
    s = jb_cadd(a5, a7); a7 = jb_csub(a5, a7); a5 = s;
    t = a5; a5.r = __dmul_rn(h, __dadd_rn(t.r, t.i)); a5.i = __dmul_rn(h, __dsub_rn(t.i, t.r));             // ROTX45
    t = a7; a7.r = __dmul_rn(h, __dsub_rn(t.i, t.r)); a7.i = __dmul_rn(h, __dsub_rn(-t.r, t.i));            // ROTX135
    a0 = jb_cadd(c[0], c[4]); a4 = jb_csub(c[0], c[4]);
    a2 = jb_cadd(c[2], c[6]); a6 = jb_csub(c[2], c[6]);
    s = jb_cadd(a0, a2); ch[0] = jb_cadd(s, a1); ch[4] = jb_csub(s, a1);
    s = jb_csub(a0, a2); ch[2] = jb_cadd(s, a3); ch[6] = jb_csub(s, a3);
    a6 = jb_rot90(a6);
    s = jb_cadd(a4, a6); ch[1] = jb_cadd(s, a5); ch[5] = jb_csub(s, a5);
    s = jb_csub(a4, a6); ch[3] = jb_cadd(s, a7); ch[7] = jb_csub(s, a7);
}
// (Not inlined, operands through memory: the single-lane path runs for a few coefficients in a thousand, and kept
// out of line it costs the calling kernels no registers.)
static __device__ __noinline__ void jb_pf_pass8(const JbC64* c, JbC64* ch) { jb_pf_pass8_body(c, ch); }

// The same transform of one 8 x 8 block by a whole warp: lanes 0..7 (and their copies 8..31) take one row each through
// the radix-8 pass in registers, the eight values of column v meet by shuffles, and a second pass down that column gives
// Re F[u][v].  X: the 64 box sums of the block (block_size 4: mean = sum / 16, exact).  Every lane of the warp must call;
// every lane gets the value.  ~300 instructions, where a lane on its own needs nine passes through local memory.
static __device__ __noinline__ double jb_pf_fft2_8x8_warp(const float* X, int u, int v) {
    const int i = threadIdx.x & 7;
    JbC64 c[8], ch[8];
    #pragma unroll
    for (int j = 0; j < 8; ++j) { c[j].r = __dmul_rn((double)X[i * 8 + j], 0.0625); c[j].i = 0.0; }
    jb_pf_pass8_body(c, ch);
    JbC64 pick = ch[0];
    #pragma unroll
    for (int k = 1; k < 8; ++k) if (k == v) pick = ch[k];
    #pragma unroll
    for (int r = 0; r < 8; ++r) {
        c[r].r = __shfl_sync(0xffffffffu, pick.r, r);
        c[r].i = __shfl_sync(0xffffffffu, pick.i, r);
    }
    jb_pf_pass8_body(c, ch);
    double y = ch[0].r;
    #pragma unroll
    for (int k = 1; k < 8; ++k) if (k == u) y = ch[k].r;
    return y;
}

// The DCT of one 8 x 8 block (block_size 4) at (u, v) by a whole warp, in the order of jb_refine_f64's DCT branch: lane i
// (and its copies) sums row i against C[v] term by term in ascending j, multiplies by C[u][i], and the eight terms are
// added in ascending i.  Every lane must call; every lane gets the value.  ~60 instructions against ~300 for one lane
// on its own with 31 waiting.  A64: the fp64 matrix (shared or global memory).
static __device__ __noinline__ double jb_dct2_8x8_warp(const float* X, int u, int v, const double* A64) {
    const int i = threadIdx.x & 7;
    double m = __dmul_rn(__dmul_rn((double)X[i * 8], 0.0625), A64[v * 8]);
    #pragma unroll
    for (int j = 1; j < 8; ++j) m = __dadd_rn(m, __dmul_rn(__dmul_rn((double)X[i * 8 + j], 0.0625), A64[v * 8 + j]));
    const double term = __dmul_rn(m, A64[u * 8 + i]);
    double y = __shfl_sync(0xffffffffu, term, 0);
    #pragma unroll
    for (int r = 1; r < 8; ++r) y = __dadd_rn(y, __shfl_sync(0xffffffffu, term, r));
    return y;
}

// X: the d x d box sums of the block (exact integers, as int or float; anything indexable by i * d + j); returns the value the reference hands to
// np.round for coefficient (u, v).  A64 / B64: the fp64 transform matrices of jb_tables.cu; recip: qrecip[u*d+v].
// (Inlined into a small __noinline__ wrapper per kernel file, so that the hot kernels see a call with few operands.)
template <typename XA>
__device__ __forceinline__ double jb_refine_f64(XA X, int u, int v, int d, int bs, int transform, int qmode,
                                             const double* A64, const double* B64, double recip) {
    // np.mean: exact integer sum / count (subsampling.py:9-11).  A power-of-two count divides exactly, so the
    // (slow) fp64 division can be a multiplication by the exact reciprocal with the same result.
    const JbBoxMean mean(bs);
    double y;
    if (transform == JB_TRANSFORM_DFT && d == 8) {
        JbC64 col[8];
        for (int i = 0; i < 8; ++i) {
            JbC64 c[8], ch[8];
            #pragma unroll
            for (int j = 0; j < 8; ++j) { c[j].r = mean((double)X[i * 8 + j]); c[j].i = 0.0; }
            jb_pf_pass8(c, ch);
            JbC64 pick = ch[0];
            #pragma unroll
            for (int k = 1; k < 8; ++k) if (k == v) pick = ch[k];
            col[i] = pick;
        }
        JbC64 out[8];
        jb_pf_pass8(col, out);
        y = out[0].r;
        #pragma unroll
        for (int k = 1; k < 8; ++k) if (k == u) y = out[k].r;
    } else if (transform == JB_TRANSFORM_DFT) {
        y = 0.0;
        for (int i = 0; i < d; ++i) {
            double mc = 0.0, ms = 0.0;
            for (int j = 0; j < d; ++j) {
                const double x = mean((double)X[i * d + j]);
                mc = __dadd_rn(mc, __dmul_rn(A64[v * d + j], x));
                ms = __dadd_rn(ms, __dmul_rn(B64[v * d + j], x));
            }
            y = __dadd_rn(y, __dsub_rn(__dmul_rn(A64[u * d + i], mc), __dmul_rn(B64[u * d + i], ms)));
        }
    } else {
        y = 0.0;
        for (int i = 0; i < d; ++i) {
            double m = __dmul_rn(mean((double)X[i * d]), A64[v * d]);
            for (int j = 1; j < d; ++j) m = __dadd_rn(m, __dmul_rn(mean((double)X[i * d + j]), A64[v * d + j]));
            const double term = __dmul_rn(m, A64[u * d + i]);
            y = i == 0 ? term : __dadd_rn(y, term);
        }
    }
    if (qmode == JB_Q_QTABLE) return __dmul_rn(y, recip);            // quantizers.py:47-49: a * (1.0 / q)
    if (qmode == JB_Q_DIVIDE) return __ddiv_rn(y, recip);            // quantizers.py:27-28: a / float(divisor)
    return y;
}
