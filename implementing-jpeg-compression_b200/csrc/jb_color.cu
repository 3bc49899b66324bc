// jb_color.cu -- RGB <-> YCbCr on the device, bit-identical to Pillow's Image.convert (the step on either side of
// the codec path: compress.py:9 `convert('YCbCr')`, decompress.py:10 `convert('RGB')`, consumed by
// pipeline/__init__.py:103-106 and produced by :120-124).  SURVEY.md section 8(f) row 1.
//
// Pillow converts with per-channel lookup tables in 6-bit fixed point; jb_color_tables.h holds tables that were fitted
// against Pillow on all 2^24 inputs (tools/derive_pil_tables.py) and reproduce it exactly:
//   y, cb, cr = (T[0][r] + T[1][g] + T[2][b]) >> 6
//   r = clip(y + (R_CR[cr] >> 6)),  g = clip(y + ((G_CB[cb] + G_CR[cr]) >> 6)),  b = clip(y + (B_CB[cb] >> 6))
// Both kernels are HBM-bound byte shuffles: interleaved pixels on one side (3 bytes per pixel, what PIL hands over),
// three planes on the other (what jb_compress_planes reads / jb_decompress_planes writes).  A thread owns 16 pixels of
// a row: three 128-bit accesses on the interleaved side, one per plane.  Forward, the three contributions of an input
// byte are packed into one 64-bit shared-memory word: y at bit 0, cb at bit 26, cr at bit 42.  The negative tables are
// biased by 8192 each -- 16384 per component in total, which vanishes when the sum is shifted by 6 and cut to a byte --
// so a pixel costs three table reads and one 64-bit three-way add, and the cb / cr results land byte-aligned in the high
// word (bits 32..39 and 48..55).
#include "jb_common.cuh"
#define JB_COLOR_TABLE_QUAL __constant__
#include "jb_color_tables.h"                // the tables live in constant memory, initialised when the module loads

#define JC_THREADS 256
#define JC_FWD_THREADS 512
#define JC_FWD_SMEM (3 * 256 * 16 * 8)     // replicated tables of the forward kernel
#define JC_CB_SHIFT 26                       // field offsets inside the packed table word
#define JC_CR_SHIFT 42
#define JC_BIAS 8192                        // added to every table with negative entries (two per chroma component)

struct JcArgs {
    const uint8_t* src;
    uint8_t* dst;
    size_t image_stride, rgb_pitch;          // interleaved side
    size_t plane_stride, plane_pitch;        // planar side: planes 3i, 3i+1, 3i+2 of image i
    int n_images, H, W;
    int vec_ok;                              // 16-byte alignment of every row on both sides
};

__device__ __forceinline__ uint32_t jc_byte(uint32_t w, int k) { return (w >> (8 * k)) & 0xFFu; }

// ---- interleaved RGB -> planes Y, Cb, Cr ----------------------------------------------------------------------------
__global__ void __launch_bounds__(JC_FWD_THREADS) jb_rgb_to_ycc_kernel(JcArgs a) {
    // per input channel and value: y | cb << 26 | cr << 42, in 16 copies -- lane l of a half-warp reads copy l, which
    // sits in its own pair of banks, so the (data dependent) table reads never conflict
    extern __shared__ unsigned long long s_rep[];
    const int cp = threadIdx.x & 15;
    for (int i = threadIdx.x >> 4; i < 768; i += JC_FWD_THREADS >> 4) {
        const int ch = i >> 8, v = i & 255;
        // cb: the R and G tables are negative, cr: the G and B tables (the +128 << 6 sits in the B tables)
        const int y = JB_RGB2YCC_Y[ch * 256 + v];
        const int cb = JB_RGB2YCC_CB[ch * 256 + v] + (ch != 2 ? JC_BIAS : 0);
        const int cr = JB_RGB2YCC_CR[ch * 256 + v] + (ch != 0 ? JC_BIAS : 0);
        s_rep[i * 16 + cp] = (unsigned long long)y | ((unsigned long long)cb << JC_CB_SHIFT) | ((unsigned long long)cr << JC_CR_SHIFT);
    }
    __syncthreads();
    const unsigned long long* t_r = s_rep + cp, * t_g = s_rep + 256 * 16 + cp, * t_b = s_rep + 512 * 16 + cp;
    const int groups = (a.W + 15) >> 4;                  // 16-pixel groups per row
    const long long total = (long long)a.n_images * a.H * groups;
    for (long long t = blockIdx.x * (long long)JC_FWD_THREADS + threadIdx.x; t < total; t += (long long)gridDim.x * JC_FWD_THREADS) {
        const int gx = (int)(t % groups);
        const long long ry = t / groups;
        const int y = (int)(ry % a.H), img = (int)(ry / a.H);
        const int x0 = gx * 16;
        const uint8_t* in = a.src + (size_t)img * a.image_stride + (size_t)y * a.rgb_pitch + (size_t)x0 * 3;
        uint8_t* out = a.dst + (size_t)(3 * img) * a.plane_stride + (size_t)y * a.plane_pitch + x0;
        const int npx = min(16, a.W - x0);
        if (!(a.vec_ok && npx == 16)) {
            // ragged row end or unaligned buffers: pixel by pixel
            for (int k = 0; k < npx; ++k) {
                const unsigned long long s = t_r[in[3 * k] * 16] + t_g[in[3 * k + 1] * 16] + t_b[in[3 * k + 2] * 16];
                out[k] = (uint8_t)((uint32_t)s >> 6);
                out[a.plane_stride + k] = (uint8_t)(s >> 32);
                out[2 * a.plane_stride + k] = (uint8_t)(s >> 48);
            }
            continue;
        }
        const uint4* in16 = (const uint4*)in;
        const uint4 q0 = __ldg(in16), q1 = __ldg(in16 + 1), q2 = __ldg(in16 + 2);
        const uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
        uint32_t oy[4], ocb[4], ocr[4];
        #pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {                 // 4 pixels = 3 words
            const uint32_t a0 = w[3 * g4], a1 = w[3 * g4 + 1], a2 = w[3 * g4 + 2];
            const uint32_t r[4] = {jc_byte(a0, 0), jc_byte(a0, 3), jc_byte(a1, 2), jc_byte(a2, 1)};
            const uint32_t g[4] = {jc_byte(a0, 1), jc_byte(a1, 0), jc_byte(a1, 3), jc_byte(a2, 2)};
            const uint32_t b[4] = {jc_byte(a0, 2), jc_byte(a1, 1), jc_byte(a2, 0), jc_byte(a2, 3)};
            uint32_t lo[4], hi[4];
            #pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned long long s = t_r[r[k] * 16] + t_g[g[k] * 16] + t_b[b[k] * 16];
                lo[k] = (uint32_t)s >> 6;                // y in the low byte
                hi[k] = (uint32_t)(s >> 32);             // cb in byte 0, cr in byte 2
            }
            oy[g4] = __byte_perm(__byte_perm(lo[0], lo[1], 0x0040), __byte_perm(lo[2], lo[3], 0x0040), 0x5410);
            ocb[g4] = __byte_perm(__byte_perm(hi[0], hi[1], 0x0040), __byte_perm(hi[2], hi[3], 0x0040), 0x5410);
            ocr[g4] = __byte_perm(__byte_perm(hi[0], hi[1], 0x0062), __byte_perm(hi[2], hi[3], 0x0062), 0x5410);
        }
        *(uint4*)out = make_uint4(oy[0], oy[1], oy[2], oy[3]);
        *(uint4*)(out + a.plane_stride) = make_uint4(ocb[0], ocb[1], ocb[2], ocb[3]);
        *(uint4*)(out + 2 * a.plane_stride) = make_uint4(ocr[0], ocr[1], ocr[2], ocr[3]);
    }
}

// ---- planes Y, Cb, Cr -> interleaved RGB ----------------------------------------------------------------------------
__global__ void __launch_bounds__(JC_THREADS) jb_ycc_to_rgb_kernel(JcArgs a) {
    __shared__ int s_cr[256], s_cb[256];                 // (R_CR >> 6) | G_CR << 16 ;  (B_CB >> 6) | G_CB << 16
    for (int v = threadIdx.x; v < 256; v += JC_THREADS) {
        s_cr[v] = ((JB_YCC2RGB_R_CR[v] >> 6) & 0xFFFF) | ((int)JB_YCC2RGB_G_CR[v] << 16);
        s_cb[v] = ((JB_YCC2RGB_B_CB[v] >> 6) & 0xFFFF) | ((int)JB_YCC2RGB_G_CB[v] << 16);
    }
    __syncthreads();
    const int groups = (a.W + 15) >> 4;
    const long long total = (long long)a.n_images * a.H * groups;
    for (long long t = blockIdx.x * (long long)JC_THREADS + threadIdx.x; t < total; t += (long long)gridDim.x * JC_THREADS) {
        const int gx = (int)(t % groups);
        const long long ry = t / groups;
        const int y = (int)(ry % a.H), img = (int)(ry / a.H);
        const int x0 = gx * 16;
        const uint8_t* in = a.src + (size_t)(3 * img) * a.plane_stride + (size_t)y * a.plane_pitch + x0;
        uint8_t* out = a.dst + (size_t)img * a.image_stride + (size_t)y * a.rgb_pitch + (size_t)x0 * 3;
        const int npx = min(16, a.W - x0);
        if (!(a.vec_ok && npx == 16)) {
            for (int k = 0; k < npx; ++k) {
                const int yy = in[k];
                const int tcb = s_cb[in[a.plane_stride + k]], tcr = s_cr[in[2 * a.plane_stride + k]];
                out[3 * k] = (uint8_t)max(0, min(255, yy + (int)(short)(tcr & 0xFFFF)));
                out[3 * k + 1] = (uint8_t)max(0, min(255, yy + (((tcr >> 16) + (tcb >> 16)) >> 6)));
                out[3 * k + 2] = (uint8_t)max(0, min(255, yy + (int)(short)(tcb & 0xFFFF)));
            }
            continue;
        }
        const uint4 q0 = __ldg((const uint4*)in), q1 = __ldg((const uint4*)(in + a.plane_stride)),
                    q2 = __ldg((const uint4*)(in + 2 * a.plane_stride));
        const uint32_t wy[4] = {q0.x, q0.y, q0.z, q0.w}, wcb[4] = {q1.x, q1.y, q1.z, q1.w}, wcr[4] = {q2.x, q2.y, q2.z, q2.w};
        uint32_t o[12];
        #pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
            uint32_t px[4];                              // r | g << 8 | b << 16 per pixel
            #pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int yy = (int)jc_byte(wy[g4], k);
                const int tcr = s_cr[jc_byte(wcr[g4], k)], tcb = s_cb[jc_byte(wcb[g4], k)];
                const int r = yy + (int)(short)(tcr & 0xFFFF);
                const int b = yy + (int)(short)(tcb & 0xFFFF);
                const int g = yy + (((tcr >> 16) + (tcb >> 16)) >> 6);
                px[k] = (uint32_t)max(0, min(255, r)) | ((uint32_t)max(0, min(255, g)) << 8) | ((uint32_t)max(0, min(255, b)) << 16);
            }
            o[3 * g4] = px[0] | (px[1] << 24);
            o[3 * g4 + 1] = (px[1] >> 8) | (px[2] << 16);
            o[3 * g4 + 2] = (px[2] >> 16) | (px[3] << 8);
        }
        uint4* o16 = (uint4*)out;
        o16[0] = make_uint4(o[0], o[1], o[2], o[3]);
        o16[1] = make_uint4(o[4], o[5], o[6], o[7]);
        o16[2] = make_uint4(o[8], o[9], o[10], o[11]);
    }
}

static int jc_launch(bool forward, const uint8_t* src, uint8_t* dst, size_t image_stride, size_t rgb_pitch,
                     size_t plane_stride, size_t plane_pitch, int n_images, int H, int W, void* stream) {
    if (n_images <= 0 || H <= 0 || W <= 0) return JB_ERR_EMPTY_ARRAY;
    if (!src || !dst || rgb_pitch < (size_t)3 * W || plane_pitch < (size_t)W) return JB_ERR_BAD_PARAM;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return JB_ERR_NO_DEVICE;
    JcArgs a;
    a.src = src; a.dst = dst;
    a.image_stride = image_stride; a.rgb_pitch = rgb_pitch; a.plane_stride = plane_stride; a.plane_pitch = plane_pitch;
    a.n_images = n_images; a.H = H; a.W = W;
    const uint8_t* rgb = forward ? src : dst;
    const uint8_t* pl = forward ? dst : src;
    a.vec_ok = (((uintptr_t)rgb & 15) == 0 && (image_stride & 15) == 0 && (rgb_pitch & 15) == 0 &&
                ((uintptr_t)pl & 15) == 0 && (plane_stride & 15) == 0 && (plane_pitch & 15) == 0) ? 1 : 0;
    const long long total = (long long)n_images * H * ((W + 15) >> 4);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaStream_t s = (cudaStream_t)stream;
    if (forward) {
        // persistent CTAs (two per SM: 96 KB of tables each), grid-stride over the 16-pixel groups
        if (cudaFuncSetAttribute(jb_rgb_to_ycc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, JC_FWD_SMEM) != cudaSuccess)
            return JB_ERR_CUDA;
        const long long want = (total + JC_FWD_THREADS - 1) / JC_FWD_THREADS;
        const long long cap = (long long)sms * 2;
        JB_LAUNCH((jb_rgb_to_ycc_kernel), (unsigned)(want < cap ? want : cap), JC_FWD_THREADS, JC_FWD_SMEM, s, a);
    } else {
        const long long want = (total + JC_THREADS - 1) / JC_THREADS;
        const long long cap = (long long)sms * 8 * 4;    // grid-stride beyond a few waves
        JB_LAUNCH((jb_ycc_to_rgb_kernel), (unsigned)(want < cap ? want : cap), JC_THREADS, 0, s, a);
    }
    return cudaGetLastError() == cudaSuccess ? JB_OK : JB_ERR_CUDA;
}

extern "C" int jb_rgb_to_ycbcr_planes(const uint8_t* d_rgb, size_t image_stride, size_t rgb_pitch, int n_images,
                                      int height, int width, uint8_t* d_planes, size_t plane_stride, size_t plane_pitch,
                                      void* stream) {
    return jc_launch(true, d_rgb, d_planes, image_stride, rgb_pitch, plane_stride, plane_pitch, n_images, height, width, stream);
}

extern "C" int jb_ycbcr_planes_to_rgb(const uint8_t* d_planes, size_t plane_stride, size_t plane_pitch, int n_images,
                                      int height, int width, uint8_t* d_rgb, size_t image_stride, size_t rgb_pitch,
                                      void* stream) {
    return jc_launch(false, d_planes, d_rgb, image_stride, rgb_pitch, plane_stride, plane_pitch, n_images, height, width, stream);
}

// =====================================================================================================================
// Container assembly for a batch of images: file_format.generate_data
// (file_format.py:86-93) writes, per image, the header followed by a little-endian u32 length and the stream of each of
// the three bands.  With the streams of image i stored back to back (planes 3i, 3i+1, 3i+2 of jb_compress_planes),
// container i starts at  i * (header_len + 12) + plane_off[3i] - plane_off[0]  -- no scan needed -- and one device to host
// copy then moves every container of the batch.
// =====================================================================================================================
struct JcPackArgs {
    const uint8_t* streams;
    const unsigned long long* plane_off;     // [3 n + 1]
    uint8_t* out;
    unsigned long long* image_off;           // [n + 1]
    unsigned long long out_cap;
    unsigned long long* status;              // [0]: error code (capacity)
    int n_images, header_len;
    uint8_t header[256];
};

__global__ void __launch_bounds__(256) jb_pack_containers_kernel(const JcPackArgs a) {
    const int img = blockIdx.x, tid = threadIdx.x;
    const unsigned long long o0 = a.plane_off[0];
    const unsigned long long s0 = a.plane_off[3 * img], s1 = a.plane_off[3 * img + 1], s2 = a.plane_off[3 * img + 2],
                             s3 = a.plane_off[3 * img + 3];
    const unsigned long long per = (unsigned long long)a.header_len + 12ull;
    const unsigned long long dst0 = (unsigned long long)img * per + (s0 - o0);
    const unsigned long long end = dst0 + per + (s3 - s0);
    if (tid == 0) {
        a.image_off[img] = dst0;
        if (img == a.n_images - 1) a.image_off[a.n_images] = end;
    }
    if (end > a.out_cap) {
        if (tid == 0) atomicMax(a.status, (unsigned long long)(-JB_ERR_OUT_CAPACITY));
        return;
    }
    uint8_t* d = a.out + dst0;
    for (int i = tid; i < a.header_len; i += 256) d[i] = a.header[i];
    d += a.header_len;
    const unsigned long long st[4] = {s0, s1, s2, s3};
    #pragma unroll
    for (int b = 0; b < 3; ++b) {
        const unsigned long long len = st[b + 1] - st[b];
        if (tid < 4) d[tid] = (uint8_t)(len >> (8 * tid));                 // struct.pack('<L', len)
        d += 4;
        const uint8_t* src = a.streams + st[b];
        // bytes up to a 16-byte boundary of the destination, then 128-bit stores assembled from unaligned loads
        const unsigned head = (unsigned)((16u - (unsigned)((uintptr_t)d & 15u)) & 15u);
        const unsigned long long h = len < head ? len : head;
        if (tid < h) d[tid] = src[tid];
        const uint8_t* s = src + h;
        uint4* d16 = (uint4*)(d + h);
        const unsigned mis = (unsigned)((uintptr_t)s & 3u);
        // (an unaligned source reads one word past each 16 bytes: keep that word inside the stream)
        const unsigned long long body = len - h, guard = mis ? 4ull : 0ull;
        const unsigned long long nvec = body > guard ? (body - guard) >> 4 : 0ull;
        const uint32_t* sw = (const uint32_t*)(s - mis);
        for (unsigned long long v = tid; v < nvec; v += 256) {
            const uint32_t w0 = __ldg(sw + 4 * v), w1 = __ldg(sw + 4 * v + 1), w2 = __ldg(sw + 4 * v + 2), w3 = __ldg(sw + 4 * v + 3);
            uint4 o = make_uint4(w0, w1, w2, w3);
            if (mis) {
                const uint32_t w4 = __ldg(sw + 4 * v + 4);
                const unsigned sh = mis * 8u;
                o = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                               __funnelshift_r(w3, w4, sh));
            }
            d16[v] = o;
        }
        const unsigned long long done = h + (nvec << 4);
        if (done + tid < len) d[done + tid] = src[done + tid];               // fewer than 16 bytes left
        d += len;
    }
}

extern "C" size_t jb_containers_max_bytes(int n_images, int header_len, size_t stream_bytes) {
    if (n_images <= 0 || header_len < 0 || header_len > 256) return 0;
    return (size_t)n_images * ((size_t)header_len + 12) + stream_bytes;
}

extern "C" int jb_pack_containers(const uint8_t* d_streams, const uint64_t* d_plane_off, int n_images, const uint8_t* header,
                                  int header_len, uint8_t* d_out, size_t out_cap, uint64_t* d_image_off, uint64_t* d_status,
                                  void* stream) {
    if (n_images <= 0) return JB_ERR_EMPTY_ARRAY;
    if (!d_streams || !d_plane_off || !header || !d_out || !d_image_off || !d_status || header_len <= 0 || header_len > 256)
        return JB_ERR_BAD_PARAM;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return JB_ERR_NO_DEVICE;
    JcPackArgs a;
    a.streams = d_streams; a.plane_off = (const unsigned long long*)d_plane_off; a.out = d_out;
    a.image_off = (unsigned long long*)d_image_off; a.out_cap = out_cap; a.status = (unsigned long long*)d_status;
    a.n_images = n_images; a.header_len = header_len;
    memcpy(a.header, header, (size_t)header_len);
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(d_status, 0, sizeof(uint64_t), s) != cudaSuccess) return JB_ERR_CUDA;
    JB_LAUNCH((jb_pack_containers_kernel), n_images, 256, 0, s, a);
    return cudaGetLastError() == cudaSuccess ? JB_OK : JB_ERR_CUDA;
}
