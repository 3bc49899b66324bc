// dram_probe_rows.cu -- write-only HBM ceiling for row segments of different widths at the codec's pitches.
// A warp owns runs of 32 consecutive 32x32-pixel blocks in raster order (a chunk, wrapping at the end of a block
// row like the decoder's chunks) and writes them either tile by tile (4 blocks = 128 B x 32 rows, the tile
// kernel's pattern) or row by row over the whole chunk (32 rows, each up to 1024 contiguous bytes: lane pair per
// block, 16 bytes per lane).  Stand-alone:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dram_probe_rows tools/dram_probe_rows.cu && ./dram_probe_rows
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// MODE 0: tile by tile (8 tiles of 4 blocks); MODE 1: row by row over the chunk, 2 stores per pixel row;
// MODE 2: row by row, 4 pixel rows of one half (512 B) then the other half
template <int MODE>
__global__ void k_chunks(uint8_t* buf, int W, int H, int planes) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int hb = W / 32, vb = H / 32, nblocks = hb * vb;
    const int cpp = (nblocks + 31) / 32;
    const size_t total = (size_t)planes * cpp;
    for (size_t c = blockIdx.x * (size_t)nw + warp; c < total; c += (size_t)gridDim.x * nw) {
        const int plane = (int)(c / cpp), blk0 = (int)(c % cpp) * 32;
        uint8_t* pl = buf + (size_t)plane * H * W;
        if (MODE == 0) {
            for (int t = 0; t < 8; ++t) {
                const int blk = blk0 + 4 * t;
                if (blk + 4 > nblocks) break;
                const int by = blk / hb, bx = blk % hb;
                if (bx + 4 > hb) continue;                       // (wrapping tiles skipped: a few per plane)
                uint8_t* base = pl + (size_t)by * 32 * W + (size_t)bx * 32;
                #pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int idx = k * 32 + lane;
                    *((uint4*)(base + (size_t)(idx >> 3) * W) + (idx & 7)) = make_uint4(idx, t, blk, plane);
                }
            }
        } else {
            // lane pair (2j, 2j+1) owns block blk0 + j (first store) and blk0 + 16 + j (second store)
            const int j = lane >> 1, half = lane & 1;
            int b0 = blk0 + j, b1 = blk0 + 16 + j;
            const bool ok0 = b0 < nblocks, ok1 = b1 < nblocks;
            const int by0 = b0 / hb, bx0 = b0 % hb, by1 = b1 / hb, bx1 = b1 % hb;
            uint8_t* p0 = pl + (size_t)by0 * 32 * W + (size_t)bx0 * 32 + half * 16;
            uint8_t* p1 = pl + (size_t)by1 * 32 * W + (size_t)bx1 * 32 + half * 16;
            const uint4 v = make_uint4(lane, blk0, plane, 9);
            if (MODE == 1) {
                #pragma unroll 4
                for (int r = 0; r < 32; ++r) {
                    if (ok0) *(uint4*)(p0 + (size_t)r * W) = v;
                    if (ok1) *(uint4*)(p1 + (size_t)r * W) = v;
                }
            } else {
                for (int m = 0; m < 8; ++m) {
                    #pragma unroll
                    for (int k = 0; k < 4; ++k) if (ok0) *(uint4*)(p0 + (size_t)(4 * m + k) * W) = v;
                    #pragma unroll
                    for (int k = 0; k < 4; ++k) if (ok1) *(uint4*)(p1 + (size_t)(4 * m + k) * W) = v;
                }
            }
        }
    }
}

int main() {
    const size_t bytes = (size_t)6400 << 20;
    uint8_t* buf;
    CHECK(cudaMalloc(&buf, bytes));
    CHECK(cudaMemset(buf, 1, bytes));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int W : {1920, 2048, 3840}) {
        const int H = 1056, planes = (int)(bytes / ((size_t)W * H));
        const double tb = (double)planes * W * H;
        for (int warps : {16, 32}) {
            for (int mode = 0; mode < 3; ++mode) {
                for (int r = 0; r < 4; ++r) {
                    if (r == 1) cudaEventRecord(e0);
                    if (mode == 0) k_chunks<0><<<148, warps * 32>>>(buf, W, H, planes);
                    else if (mode == 1) k_chunks<1><<<148, warps * 32>>>(buf, W, H, planes);
                    else k_chunks<2><<<148, warps * 32>>>(buf, W, H, planes);
                }
                cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
                printf("pitch %4d, %2d warps/SM, %s : %7.1f GB/s\n", W, warps,
                       mode == 0 ? "tiles 32 rows x 128 B         " : mode == 1 ? "chunk rows 2 x 512 B per row  " : "chunk rows 4 rows x 512 B x 2 ",
                       3.0 * tb / ms * 1e-6);
            }
        }
    }
    return 0;
}
