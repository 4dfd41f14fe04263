// tools/tex_probe.cu — can the voxel fetch of the cache-resident (z slab) path leave the LSU's one-line-per-clock tag stage?
// Dependent random 4-byte gathers from a slab-sized table (1.44 MB = C2's 600 x 600 words, 4 MB = C5's) through
//   ldg   : ld.global.nc.L2::64B (what the walk kernels issue)
//   tex1d : tex1Dfetch on a linear-memory texture object (TEX pipe)
//   tex2d : tex2D point sampling on a pitched 2-D texture (TEX pipe, 2-D locality of the tag)
//   lds   : a 64 KB window of the table in shared memory (upper bound of what a software cache could give)
// with the memory-level parallelism of the walk (1 dependent gather per thread, 2048 threads per SM).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/tex_probe tools/tex_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int PATH>
__global__ void __launch_bounds__(256) probe(const uint32_t *tab, cudaTextureObject_t t1, cudaTextureObject_t t2, uint32_t nx, uint32_t ny, int iters, uint32_t *sink)
{
    extern __shared__ uint32_t win[];
    const uint32_t n = nx * ny;
    if (PATH == 3) {
        for (uint32_t i = threadIdx.x; i < 16384u; i += blockDim.x) win[i] = tab[i];
        __syncthreads();
    }
    uint32_t s = hash32((blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u);
    uint32_t acc = 0;
    for (int it = 0; it < iters; it++) {
        s = hash32(s + 0x9e3779b9u);
        const uint32_t a = (uint32_t)(((uint64_t)s * n) >> 32);
        uint32_t v;
        if (PATH == 0) asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(tab + a));
        else if (PATH == 1) v = tex1Dfetch<unsigned int>(t1, (int)a);
        else if (PATH == 2) v = tex2D<unsigned int>(t2, (float)(a % ny) + 0.5f, (float)(a / ny) + 0.5f);
        else v = win[a & 16383u];
        acc += v;
        s ^= (v & 1u); // the next address waits for this word, like the walk's accept / reject
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int PATH>
double run(const uint32_t *tab, cudaTextureObject_t t1, cudaTextureObject_t t2, uint32_t nx, uint32_t ny, int sm, int iters, uint32_t *sink)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int grid = 8 * sm;
    const size_t sh = PATH == 3 ? 65536 : 0;
    if (sh) CK(cudaFuncSetAttribute(probe<PATH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    probe<PATH><<<grid, 256, sh>>>(tab, t1, t2, nx, ny, iters / 8, sink);
    CK(cudaEventRecord(a));
    probe<PATH><<<grid, 256, sh>>>(tab, t1, t2, nx, ny, iters, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    return (double)grid * 256 * iters / (ms * 1e-3);
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sm = prop.multiProcessorCount;
    printf("path,table,gathers_per_s,per_clk_per_sm\n");
    for (uint32_t nside : {600u, 1000u}) {
        const uint32_t nx = nside, ny = nside, n = nx * ny;
        std::vector<uint32_t> h(n);
        for (uint32_t i = 0; i < n; i++) h[i] = i * 2654435761u;
        uint32_t *tab, *sink;
        CK(cudaMalloc(&tab, n * 4)); CK(cudaMalloc(&sink, 4));
        CK(cudaMemcpy(tab, h.data(), n * 4, cudaMemcpyHostToDevice));
        cudaResourceDesc rd{}; cudaTextureDesc td{};
        rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = tab; rd.res.linear.desc = cudaCreateChannelDesc<unsigned int>(); rd.res.linear.sizeInBytes = (size_t)n * 4;
        td.readMode = cudaReadModeElementType; td.filterMode = cudaFilterModePoint; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        cudaTextureObject_t t1, t2;
        CK(cudaCreateTextureObject(&t1, &rd, &td, nullptr));
        cudaArray_t arr;
        cudaChannelFormatDesc cd = cudaCreateChannelDesc<unsigned int>();
        CK(cudaMallocArray(&arr, &cd, ny, nx)); // block-linear layout: 2-D locality
        CK(cudaMemcpy2DToArray(arr, 0, 0, h.data(), ny * 4, ny * 4, nx, cudaMemcpyHostToDevice));
        cudaResourceDesc ra{}; ra.resType = cudaResourceTypeArray; ra.res.array.array = arr;
        CK(cudaCreateTextureObject(&t2, &ra, &td, nullptr));
        const int iters = 4096;
        const double clk = prop.clockRate * 1e3;
        const char *names[4] = {"ldg", "tex1d", "tex2d_array", "lds_64KB_window"};
        double r[4] = {run<0>(tab, t1, t2, nx, ny, sm, iters, sink), run<1>(tab, t1, t2, nx, ny, sm, iters, sink), run<2>(tab, t1, t2, nx, ny, sm, iters, sink),
                       run<3>(tab, t1, t2, nx, ny, sm, iters, sink)};
        for (int p = 0; p < 4; p++) printf("%s,%ux%u,%.4g,%.3f\n", names[p], nx, ny, r[p], r[p] / clk / sm);
        CK(cudaDestroyTextureObject(t1)); CK(cudaDestroyTextureObject(t2)); CK(cudaFreeArray(arr)); CK(cudaFree(tab)); CK(cudaFree(sink));
    }
    return 0;
}
