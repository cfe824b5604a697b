// Micro-probe for the blob hand-over question (profiles/r01_stage_notes.md): how fast does HBM take pure writes, and
// does a consumer that reads freshly written lines out of L2 and then issues discard.global.L2 keep them from ever
// being written back?  Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_bin/l2_probe scripts/l2_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__global__ void fill_kernel(uint4* p, size_t n, unsigned v, int streaming)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += step) {
        uint4 w = make_uint4(v, (unsigned)i, v, (unsigned)i);
        if (streaming) __stcs(p + i, w); else p[i] = w;
    }
}

__global__ void consume_kernel(const uint4* p, size_t n, int discard, unsigned long long* out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (; i < n; i += step) {                      // n is a multiple of the grid size: no divergence in the loop
        uint4 w = p[i];
        unsigned s = w.x + w.y + w.z + w.w;
        s += __shfl_xor_sync(0xffffffffu, s, 1);    // the 8 lanes of a 128-byte line have all consumed their loads
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        acc += s;
        if (discard && (threadIdx.x & 7) == 0)
            asm volatile("discard.global.L2 [%0], 128;" :: "l"(p + i) : "memory");
    }
    if (acc == 0x1234567ull) *out = acc;
}

int main(int argc, char** argv)
{
    const size_t buf_bytes = (argc > 1 ? atol(argv[1]) : 32) << 20;   // per hand-over buffer (MiB)
    const int nbuf = argc > 2 ? atoi(argv[2]) : 24;
    const size_t n = buf_bytes / 16;
    uint4* pool; unsigned long long* out;
    CK(cudaMalloc(&pool, buf_bytes * nbuf)); CK(cudaMalloc(&out, 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grid = 148 * 8, block = 256;
    float ms;

    for (int streaming = 0; streaming < 2; ++streaming) {            // 1. pure write bandwidth over the whole pool
        fill_kernel<<<grid, block>>>(pool, n * nbuf, 1, streaming);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int r = 0; r < 4; ++r) fill_kernel<<<grid, block>>>(pool, n * nbuf, 2 + r, streaming);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("fill %s: %.1f GB/s (%zu MiB x4)\n", streaming ? "st.cs" : "st   ", 4.0 * buf_bytes * nbuf / ms * 1e-6, (buf_bytes * nbuf) >> 20);
    }
    for (int discard = 0; discard < 2; ++discard) {                  // 2. write a buffer, read it back at once, next buffer
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int r = 0; r < 2; ++r)
            for (int b = 0; b < nbuf; ++b) {
                fill_kernel<<<grid, block>>>(pool + (size_t)b * n, n, 7 + r, 0);
                consume_kernel<<<grid, block>>>(pool + (size_t)b * n, n, discard, out);
            }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("hand-over discard=%d: %.3f ms for %d x %zu MiB  (%.1f GB/s of payload)\n", discard, ms, 2 * nbuf, buf_bytes >> 20,
               2.0 * nbuf * buf_bytes / ms * 1e-6);
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
