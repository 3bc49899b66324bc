// jb_stage_f64.cu -- parity hooks that expose the float64 intermediates of the compress direction, stage by stage,
// the way the reference's pipeline holds them between its stages (pipeline/base.py:42-72: every stage has an
// `execute` whose result the next stage receives):
//
//   JB_F64_SAMPLES      Padding + SubSampling + DCTPadding (padding.py:8-12, subsampling.py:9-11, dct_padding.py:8-9):
//                       the (vb d) x (hb d) float64 array of box means BasisChange receives
//   JB_F64_TRANSFORM    + BasisChange.execute (basis_change.py:11-26, transforms.py:46-58): per block d x d values
//                       in natural (u, v) order -- the real part for the DFT (the imaginary part never reaches the
//                       stream: run_length_encoding.py:16-17)
//   JB_F64_PREROUNDING  + the quantiser's scaling (quantizers.py:5-6, 16-17, 27-28, 47-49), the value np.round receives
//
// The arithmetic is jb_refine.cuh -- the float64 re-evaluation the fused kernels fall back on when their fp32 value
// lies within its error bound of a rounding tie -- applied to every coefficient instead of a few in a thousand, so a
// test that compares these arrays bit for bit with the oracle pins exactly the code that decides ties.  Not a fast
// path: one CTA per block, one thread per coefficient, box sums staged in shared memory.
#include <string.h>

#include "jb_common.cuh"
#include "jb_refine.cuh"

struct JbF64Args {
    JbGeom g;
    JbTables t;
    const uint8_t* planes;
    size_t plane_stride, row_pitch;
    int n_planes;
    int which;
    double* out;
};

// box sum of sample (si, sj) of the subsampled, dct-padded plane: the reference replicates the edge twice -- pixels up
// to a multiple of block_size (padding.py:8-12), then box means up to a multiple of dct_size (dct_padding.py:8-9)
__device__ __forceinline__ unsigned jb_f64_box_sum(const JbF64Args& a, const uint8_t* plane, int si, int sj) {
    const JbGeom& g = a.g;
    const int ci = jb_min(si, g.H1 - 1), cj = jb_min(sj, g.W1 - 1);
    unsigned sum = 0;
    for (int r = 0; r < g.bs; ++r) {
        const uint8_t* row = plane + (size_t)jb_min(ci * g.bs + r, g.H - 1) * a.row_pitch;
        for (int c = 0; c < g.bs; ++c) sum += row[jb_min(cj * g.bs + c, g.W - 1)];
    }
    return sum;
}

__global__ void __launch_bounds__(256) jb_stage_f64_kernel(const JbF64Args a) {
    __shared__ float s_x[JB_MAX_DCT_SIZE * JB_MAX_DCT_SIZE];             // box sums of the block (integers below 2^24)
    const JbGeom& g = a.g;
    const int d = g.d, n = g.n;
    const int plane = blockIdx.y;
    const int blk = blockIdx.x;
    const int by = blk / g.hb, bx = blk - by * g.hb;
    const uint8_t* src = a.planes + (size_t)plane * a.plane_stride;
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const int i = idx / d, j = idx - i * d;
        s_x[idx] = (float)jb_f64_box_sum(a, src, by * d + i, bx * d + j);
    }
    __syncthreads();
    const JbBoxMean mean(g.bs);
    if (a.which == JB_F64_SAMPLES) {
        const size_t w2 = (size_t)g.hb * d;
        double* out = a.out + (size_t)plane * ((size_t)g.vb * d) * w2;
        for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
            const int i = idx / d, j = idx - i * d;
            out[((size_t)by * d + i) * w2 + (size_t)bx * d + j] = mean((double)s_x[idx]);
        }
        return;
    }
    const int qmode = a.which == JB_F64_PREROUNDING ? g.qmode : JB_Q_NONE;
    double* out = a.out + ((size_t)plane * g.nblocks + blk) * n;
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const int u = idx / d, v = idx - u * d;
        out[idx] = jb_refine_f64((const float*)s_x, u, v, d, g.bs, g.transform, qmode, a.t.fA64, a.t.fB64, a.t.qrecip[idx]);
    }
}

cudaError_t jb_launch_stage_f64(const JbGeom& g, const JbTables& t, const uint8_t* d_planes, size_t plane_stride,
                                size_t row_pitch, int n_planes, int which, double* d_out, cudaStream_t s) {
    JbF64Args a;
    memset(&a, 0, sizeof(a));
    a.g = g;
    a.t = t;
    a.planes = d_planes; a.plane_stride = plane_stride; a.row_pitch = row_pitch;
    a.n_planes = n_planes; a.which = which; a.out = d_out;
    JB_LAUNCH((jb_stage_f64_kernel), dim3((unsigned)g.nblocks, (unsigned)n_planes), 256, 0, s, a);
    return cudaGetLastError();
}
