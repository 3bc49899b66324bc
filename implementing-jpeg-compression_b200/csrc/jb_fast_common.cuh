// jb_fast_common.cuh -- pieces shared by the specialised 8x8 kernels (jb_forward_fast.cu,
// jb_inverse_fast.cu): TMA / mbarrier PTX wrappers and the 8-point transforms.
#pragma once
#include <cuda.h>
#include <stdint.h>

#define FF_TILE_BYTES 4096
#define FF_COEF_W 36        // words per coefficient row: 128 B of int16 + 16 B pad (conflict-free LDS.128)
#define FF_BLK_W 72         // words per block in the 8x8 scratch arrays (64 + 8: bank rotation)

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ff_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ff_mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(ff_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ff_mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(ff_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ff_mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(ff_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool ff_mbar_try_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(ff_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void ff_tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z,
                                               unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
                 "[%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(ff_smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(ff_smem_u32(bar)) : "memory");
}
// 1-D bulk copy global -> shared (16-byte aligned on both sides, size a multiple of 16), completion on an mbarrier
__device__ __forceinline__ void ff_bulk_load_1d(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(ff_smem_u32(dst)), "l"(src), "r"(bytes), "r"(ff_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ff_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- 8-point transforms ---------------------------------------------------------------------
// y[v] = sum_j x[j] cos(pi (2j+1) v / 16)   (transforms.py:4-11, un-normalised), even/odd split
__device__ __forceinline__ void ff_dct8(const float (&x)[8], float (&y)[8]) {
    const float c1 = 0.98078528040323043f, c2 = 0.92387953251128674f, c3 = 0.83146961230254524f,
                c4 = 0.70710678118654752f, c5 = 0.55557023301960222f, c6 = 0.38268343236508977f,
                c7 = 0.19509032201612827f;
    const float s0 = x[0] + x[7], s1 = x[1] + x[6], s2 = x[2] + x[5], s3 = x[3] + x[4];
    const float d0 = x[0] - x[7], d1 = x[1] - x[6], d2 = x[2] - x[5], d3 = x[3] - x[4];
    const float e0 = s0 + s3, e1 = s1 + s2, e2 = s0 - s3, e3 = s1 - s2;
    y[0] = e0 + e1;
    y[4] = c4 * (e0 - e1);
    y[2] = fmaf(c2, e2, c6 * e3);
    y[6] = fmaf(c6, e2, -c2 * e3);
    y[1] = fmaf(c1, d0, fmaf(c3, d1, fmaf(c5, d2, c7 * d3)));
    y[3] = fmaf(c3, d0, fmaf(-c7, d1, fmaf(-c1, d2, -c5 * d3)));
    y[5] = fmaf(c5, d0, fmaf(-c1, d1, fmaf(c7, d2, c3 * d3)));
    y[7] = fmaf(c7, d0, fmaf(-c5, d1, fmaf(c3, d2, -c1 * d3)));
}

// Real 8-point DFT, packed: r[0..4] = sum_j x[j] cos(2 pi v j / 8) for v = 0..4,
// r[5..7] = sum_j x[j] sin(2 pi v j / 8) for v = 1..3 (the other outputs follow by symmetry).
__device__ __forceinline__ void ff_rdft8(const float (&x)[8], float (&r)[8]) {
    const float h = 0.70710678118654752f;
    const float a0 = x[0] + x[4], a1 = x[0] - x[4], a2 = x[2] + x[6], a3 = x[2] - x[6];
    const float b0 = x[1] + x[5], b1 = x[1] - x[5], b2 = x[3] + x[7], b3 = x[3] - x[7];
    const float p = a0 + a2, q = b0 + b2, m = b1 - b3, n = b1 + b3;
    r[0] = p + q;
    r[4] = p - q;
    r[2] = a0 - a2;
    r[6] = b0 - b2;                 // sin, v = 2
    r[1] = fmaf(h, m, a1);
    r[3] = fmaf(-h, m, a1);
    r[5] = fmaf(h, n, a3);          // sin, v = 1
    r[7] = fmaf(h, n, -a3);         // sin, v = 3
}


// x[n] = sum_k z[k] cos(pi (2n+1) k / 16): the transpose of ff_dct8 (inverse up to the
// per-frequency scale 1/|c_k|^2 of transforms.py:14-26, which the caller folds into z)
__device__ __forceinline__ void ff_idct8(const float (&z)[8], float (&x)[8]) {
    const float c1 = 0.98078528040323043f, c2 = 0.92387953251128674f, c3 = 0.83146961230254524f,
                c4 = 0.70710678118654752f, c5 = 0.55557023301960222f, c6 = 0.38268343236508977f,
                c7 = 0.19509032201612827f;
    const float p = fmaf(c4, z[4], z[0]), q = fmaf(-c4, z[4], z[0]);
    const float r = fmaf(c2, z[2], c6 * z[6]), s = fmaf(c6, z[2], -c2 * z[6]);
    const float e0 = p + r, e3 = p - r, e1 = q + s, e2 = q - s;
    const float o0 = fmaf(c1, z[1], fmaf(c3, z[3], fmaf(c5, z[5], c7 * z[7])));
    const float o1 = fmaf(c3, z[1], fmaf(-c7, z[3], fmaf(-c1, z[5], -c5 * z[7])));
    const float o2 = fmaf(c5, z[1], fmaf(-c1, z[3], fmaf(c7, z[5], c3 * z[7])));
    const float o3 = fmaf(c7, z[1], fmaf(-c5, z[3], fmaf(c3, z[5], -c1 * z[7])));
    x[0] = e0 + o0; x[7] = e0 - o0;
    x[1] = e1 + o1; x[6] = e1 - o1;
    x[2] = e2 + o2; x[5] = e2 - o2;
    x[3] = e3 + o3; x[4] = e3 - o3;
}

// Column stage of the 2-D real-part DFT.  Lane c (0..7) of a block holds, over the 8 rows,
// one packed column of the row stage: cos columns for c = 0..4 (frequency c), sin columns
// for c = 5..7 (frequency c - 4).  Returns y[u] = Re F[u][v] for v = (c < 5 ? c : 12 - c):
//   Re F[u][v] = CT(cos col v)[u] - ST(sin col v)[u],  Re F[u][8-v] = CT + ST.
// The partner column sits 4 lanes-of-c away, i.e. lane ^ 16 in the (c * 4 + b) lane layout.
__device__ __forceinline__ void ff_dft_column_stage(const float (&col)[8], int c, float (&y)[8]) {
    float t[8];
    ff_rdft8(col, t);
    // A cos lane holds CT[u] = t[min(u, 8 - u)], a sin lane ST[u] = (0, t5, t6, t7, 0, -t7, -t6, -t5): five distinct
    // values each, so five exchanges do.  With (A, S) = (cos part, sin part) of this lane's output column:
    //   y[k] = A[k] + f S[k], y[8-k] = A[k] - f S[k] (k = 1..3), y[0] = A[0], y[4] = A[4];
    //   f = -1 on the cos lanes 1..3 (column v = c), +1 on the sin lanes (column 8 - v), 0 for c = 0 and c = 4 (no sin part).
    const bool sin_lane = c >= 5;
    const float f = sin_lane ? 1.f : ((c & 3) ? -1.f : 0.f);
    float m[5], o[5];
    m[0] = sin_lane ? 0.f : t[0];
    m[1] = sin_lane ? t[5] : t[1];
    m[2] = sin_lane ? t[6] : t[2];
    m[3] = sin_lane ? t[7] : t[3];
    m[4] = sin_lane ? 0.f : t[4];
    #pragma unroll
    for (int k = 0; k < 5; ++k) o[k] = __shfl_xor_sync(0xffffffffu, m[k], 16);
    y[0] = sin_lane ? o[0] : m[0];
    y[4] = sin_lane ? o[4] : m[4];
    #pragma unroll
    for (int k = 1; k < 4; ++k) {
        const float A = sin_lane ? o[k] : m[k], S = sin_lane ? m[k] : o[k];
        y[k] = fmaf(f, S, A);
        y[8 - k] = fmaf(-f, S, A);
    }
}

// The output planes are written once and not read again by this call: evict-first keeps the 126 MB L2 from holding
// them back (measured: about -0.02 ms on the 1024-image decompress, evict-last +0.03 ms; the same hint on the
// compress kernel's tile loads made no difference).
__device__ __forceinline__ void ff_tma_store_3d(const CUtensorMap* map, int x, int y, int z, const void* src) {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2, %3}], [%4], %5;"
                 :: "l"(map), "r"(x), "r"(y), "r"(z), "r"(ff_smem_u32(src)), "l"(pol) : "memory");
}
// 1-D bulk copy shared -> global (16-byte aligned on both sides, size a multiple of 16), bulk-group completion
__device__ __forceinline__ void ff_bulk_store_1d(void* dst, const void* src, uint32_t bytes) {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 :: "l"(dst), "r"(ff_smem_u32(src)), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void ff_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ff_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }

// 3-D uint8 tensor map over a batch of planes (x = byte in row, y = row, z = plane), box box_w x box_h x 1
bool jb_make_plane_tensor_map(CUtensorMap* map, const void* base, int W, int H, int n_planes, size_t row_pitch,
                              size_t plane_stride, int box_w = 128, int box_h = 32);
