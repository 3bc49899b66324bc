// dram_probe.cu -- one-directional HBM ceilings on this GPU: pure read and pure write with plain
// 128-bit accesses and with bulk (TMA) copies through shared memory.  Stand-alone:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dram_probe tools/dram_probe.cu && ./dram_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstring>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_write(uint4* dst, size_t n) {
    const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}

__global__ void k_read(const uint4* src, size_t n, unsigned* sink) {
    unsigned acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(src + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

// each warp owns a 4 KB slot of shared memory and streams it out with cp.async.bulk (two in flight)
template <int SLOT_BYTES>
__global__ void k_bulk_write(uint8_t* dst, size_t nslots) {
    extern __shared__ __align__(128) uint8_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    uint8_t* mine = sm + (size_t)warp * 2 * SLOT_BYTES;
    for (int i = lane; i < 2 * SLOT_BYTES / 4; i += 32) ((uint32_t*)mine)[i] = i;
    __syncwarp();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    unsigned seq = 0;
    for (size_t s = blockIdx.x * (size_t)nw + warp; s < nslots; s += (size_t)gridDim.x * nw, ++seq) {
        if (lane == 0) {
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(mine + (seq & 1) * SLOT_BYTES);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(dst + s * SLOT_BYTES), "r"(sa), "r"(SLOT_BYTES) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// the codec's access pattern: planes of H x W bytes; a warp owns chunks of 8 tiles side by side, a tile is
// 32 rows x 128 bytes at the plane's row pitch; chunks are dealt round-robin to the warps of the grid
template <bool WRITE>
__global__ void k_tiles(uint8_t* buf, int W, int H, int planes, unsigned* sink) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int tiles_x = W / 128, tiles_y = H / 32;
    const int chunks_per_plane = tiles_x * tiles_y / 8;
    const size_t total = (size_t)planes * chunks_per_plane;
    unsigned acc = 0;
    for (size_t c = blockIdx.x * (size_t)nw + warp; c < total; c += (size_t)gridDim.x * nw) {
        const int plane = (int)(c / chunks_per_plane), cc = (int)(c % chunks_per_plane);
        for (int t = 0; t < 8; ++t) {
            const int tile = cc * 8 + t, ty = tile / tiles_x, tx = tile % tiles_x;
            uint8_t* base = buf + ((size_t)plane * H + (size_t)ty * 32) * W + (size_t)tx * 128;
            #pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int idx = k * 32 + lane;
                uint4* p = (uint4*)(base + (size_t)(idx >> 3) * W) + (idx & 7);
                if (WRITE) *p = make_uint4(idx, t, cc, plane);
                else { const uint4 v = __ldg(p); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
            }
        }
    }
    if (!WRITE && acc == 0x12345678u) *sink = acc;
}

// write-only, general tile shape: a warp writes tiles of ROWS x COLS bytes (COLS a multiple of 128), tiles of a
// block row side by side; ROWS_BAND = rows of a band of tiles (32 for the codec)
__global__ void k_tiles_shape(uint8_t* buf, int W, int H, int planes, int rows, int cols) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int tiles_x = W / cols, tiles_y = H / rows;
    const size_t total = (size_t)planes * tiles_x * tiles_y;
    const int vec_per_row = cols / 16, vecs = rows * vec_per_row;
    for (size_t c = blockIdx.x * (size_t)nw + warp; c < total; c += (size_t)gridDim.x * nw) {
        const int plane = (int)(c / ((size_t)tiles_x * tiles_y)), tile = (int)(c % ((size_t)tiles_x * tiles_y));
        const int ty = tile / tiles_x, tx = tile % tiles_x;
        uint8_t* base = buf + ((size_t)plane * H + (size_t)ty * rows) * W + (size_t)tx * cols;
        for (int idx = lane; idx < vecs; idx += 32)
            *((uint4*)(base + (size_t)(idx / vec_per_row) * W) + (idx % vec_per_row)) = make_uint4(idx, tile, plane, 7);
    }
}

// TMA tensor stores / loads with the codec's geometry: each warp owns two 4 KB slots (box_w x box_h bytes), tiles of a
// block row side by side, dealt round-robin to warps; `depth` stores may be in flight per warp (1 or 2)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int DEPTH>
__global__ void k_tma_store(const __grid_constant__ CUtensorMap map, int W, int H, int planes, int box_w, int box_h) {
    extern __shared__ __align__(128) uint8_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    uint8_t* mine = sm + (size_t)warp * 2 * 4096;
    for (int i = lane; i < 2 * 4096 / 4; i += 32) ((uint32_t*)mine)[i] = i;
    __syncwarp();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int tiles_x = W / box_w, tiles_y = H / box_h;
    const size_t total = (size_t)planes * tiles_x * tiles_y;
    unsigned seq = 0;
    for (size_t c = blockIdx.x * (size_t)nw + warp; c < total; c += (size_t)gridDim.x * nw, ++seq) {
        const int plane = (int)(c / ((size_t)tiles_x * tiles_y)), tile = (int)(c % ((size_t)tiles_x * tiles_y));
        const int ty = tile / tiles_x, tx = tile % tiles_x;
        if (lane == 0) {
            if (DEPTH == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(mine + (seq & 1) * 4096);
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                         :: "l"(&map), "r"(tx * box_w), "r"(ty * box_h), "r"(plane), "r"(sa) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    const size_t bytes = (size_t)6400 << 20;
    uint8_t* buf; unsigned* sink;
    CHECK(cudaMalloc(&buf, bytes));
    CHECK(cudaMalloc(&sink, 4));
    CHECK(cudaMemset(buf, 1, bytes));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const int ctas[] = {148 * 2, 148 * 4, 148 * 8, 148 * 16, 148 * 32};
    for (int threads : {256, 512, 1024}) {
        for (int g : ctas) {
            if ((long)g * threads > 148L * 2048 * 4) continue;
            k_write<<<g, threads>>>((uint4*)buf, bytes / 16);
            cudaEventRecord(e0);
            for (int r = 0; r < 3; ++r) k_write<<<g, threads>>>((uint4*)buf, bytes / 16);
            cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
            printf("write stg128  grid %5d x %4d : %7.1f GB/s\n", g, threads, 3.0 * bytes / ms * 1e-6);
            k_read<<<g, threads>>>((const uint4*)buf, bytes / 16, sink);
            cudaEventRecord(e0);
            for (int r = 0; r < 3; ++r) k_read<<<g, threads>>>((const uint4*)buf, bytes / 16, sink);
            cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
            printf("read  ldg128  grid %5d x %4d : %7.1f GB/s\n", g, threads, 3.0 * bytes / ms * 1e-6);
        }
    }
    for (int warps : {8, 14, 24}) {
        const size_t smem = (size_t)warps * 2 * 4096;
        CHECK(cudaFuncSetAttribute(k_bulk_write<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int per_sm : {1, 2}) {
            k_bulk_write<4096><<<148 * per_sm, warps * 32, smem>>>(buf, bytes / 4096);
            cudaEventRecord(e0);
            for (int r = 0; r < 3; ++r) k_bulk_write<4096><<<148 * per_sm, warps * 32, smem>>>(buf, bytes / 4096);
            cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
            printf("write bulk4K  %2d warps x %d CTA/SM : %7.1f GB/s\n", warps, per_sm, 3.0 * bytes / ms * 1e-6);
        }
    }
    {
        const int W = 1920, H = 1056, planes = (int)(bytes / ((size_t)W * H));
        const double tb = (double)planes * W * H;
        for (int warps : {8, 14, 16, 32}) {
            for (int per_sm : {1, 2}) {
                if (warps * per_sm > 64) continue;
                k_tiles<true><<<148 * per_sm, warps * 32>>>(buf, W, H, planes, sink);
                cudaEventRecord(e0);
                for (int r = 0; r < 3; ++r) k_tiles<true><<<148 * per_sm, warps * 32>>>(buf, W, H, planes, sink);
                cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
                printf("write tiles 32x128 pitch 1920  %2d warps x %d CTA/SM : %7.1f GB/s\n", warps, per_sm, 3.0 * tb / ms * 1e-6);
                k_tiles<false><<<148 * per_sm, warps * 32>>>(buf, W, H, planes, sink);
                cudaEventRecord(e0);
                for (int r = 0; r < 3; ++r) k_tiles<false><<<148 * per_sm, warps * 32>>>(buf, W, H, planes, sink);
                cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
                printf("read  tiles 32x128 pitch 1920  %2d warps x %d CTA/SM : %7.1f GB/s\n", warps, per_sm, 3.0 * tb / ms * 1e-6);
            }
        }
    }
    {
        const int shapes[][2] = {{32, 128}, {16, 256}, {8, 512}, {4, 1024}, {32, 256}, {32, 512}, {32, 1024}, {32, 1920}, {1, 1920}};
        for (int W : {1920, 2048}) {
            const int H = 1056, planes = (int)(bytes / ((size_t)W * H));
            for (auto& sh : shapes) {
                if (sh[1] > W) continue;             // (a row may end with a strip no tile covers)
                const double tb = (double)planes * (W / sh[1] * sh[1]) * H;
                k_tiles_shape<<<148 * 2, 14 * 32>>>(buf, W, H, planes, sh[0], sh[1]);
                cudaEventRecord(e0);
                for (int r = 0; r < 3; ++r) k_tiles_shape<<<148 * 2, 14 * 32>>>(buf, W, H, planes, sh[0], sh[1]);
                cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
                printf("write tiles %2d rows x %4d B, pitch %d : %7.1f GB/s\n", sh[0], sh[1], W, 3.0 * tb / ms * 1e-6);
            }
        }
    }
    {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres));
        PFN_encodeTiled enc = (PFN_encodeTiled)fp;
        const int W = 1920, H = 1056, planes = (int)(bytes / ((size_t)W * H));
        const int boxes[][2] = {{128, 32}, {256, 16}, {128, 16}, {256, 32}};
        for (auto& bx : boxes) {
            CUtensorMap map;
            memset(&map, 0, sizeof(map));
            cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
            cuuint64_t strides[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
            cuuint32_t box[3] = {(cuuint32_t)bx[0], (cuuint32_t)bx[1], 1}, estr[3] = {1, 1, 1};
            if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
                printf("encode failed\n");
                continue;
            }
            if (bx[0] * bx[1] > 8192) continue;
            const double tb = (double)planes * (W / bx[0] * bx[0]) * (H / bx[1] * bx[1]);
            const size_t smem = (size_t)14 * 2 * 4096;
            CHECK(cudaFuncSetAttribute(k_tma_store<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CHECK(cudaFuncSetAttribute(k_tma_store<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            for (int depth = 1; depth <= 2; ++depth) {
                if (bx[0] * bx[1] > 4096 && depth == 2) continue;
                for (int r = 0; r < 4; ++r) {
                    if (r == 1) cudaEventRecord(e0);
                    if (depth == 2) k_tma_store<2><<<148, 14 * 32, smem>>>(map, W, H, planes, bx[0], bx[1]);
                    else k_tma_store<1><<<148, 14 * 32, smem>>>(map, W, H, planes, bx[0], bx[1]);
                }
                cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
                printf("write TMA tensor box %3d x %2d, %d in flight per warp, 14 warps : %7.1f GB/s\n", bx[0], bx[1], depth, 3.0 * tb / ms * 1e-6);
            }
        }
    }
    // cudaMemset / cudaMemcpy for reference
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r) cudaMemsetAsync(buf, 7, bytes);
    cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemset                      : %7.1f GB/s\n", 3.0 * bytes / ms * 1e-6);
    return 0;
}
