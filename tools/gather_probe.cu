// tools/gather_probe.cu — measures the B200's random 4-byte gather rate (the roofline of the walk kernel's voxel fetch).
//
// The walk kernel's only per-step memory traffic is one 4-byte voxel word at a data-independent pseudo-random address
// (SURVEY §8d).  What bounds it is not HBM bytes/s but sectors/s: every gather moves one 32 B sector (L1<-L2) and one
// L2 fetch granule (L2<-HBM).  This probe issues exactly that access pattern with NO other work, for
//   * table sizes that are L2-resident (64 MB) and HBM-resident (864 MB = the 600^3 packed phantom, 4 GB = 1000^3),
//   * load flavours (ld.global.nc / .cg / .nc.L1::no_allocate),
//   * memory-level parallelism per thread (independent gathers in flight),
//   * spread: addresses uniform in a window of W words around a per-thread centre (W = whole table => fully random).
// Output: one CSV line per configuration: gathers/s and the equivalent sector GB/s.
//
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/gather_probe tools/gather_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{ // lowbias32
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
template <int FLAV> __device__ __forceinline__ uint32_t ld(const uint32_t *p)
{
    uint32_t v;
    if (FLAV == 0) asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (FLAV == 1) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (FLAV == 2) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (FLAV == 3) asm volatile("ld.global.nc.L1::evict_first.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (FLAV == 4) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (FLAV == 5) asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (FLAV == 6) { uint32_t b; asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(b) : "l"(p)); v = b; }
    else if (FLAV == 7) asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (FLAV == 8) {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    } else asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

template <int FLAV, int MLP>
__global__ void __launch_bounds__(256) probe(const uint32_t *tab, uint32_t n_words, uint32_t window, int iters, uint32_t *sink)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s = hash32(tid * 2654435761u + 12345u);
    const uint32_t centre = hash32(s) % n_words;
    uint32_t acc = 0;
    for (int it = 0; it < iters; it++) {
        uint32_t v[MLP];
#pragma unroll
        for (int m = 0; m < MLP; m++) {
            s = hash32(s + 0x9e3779b9u);
            uint32_t a = window >= n_words ? (uint32_t)(((uint64_t)s * n_words) >> 32)
                                           : (centre + (uint32_t)(((uint64_t)s * window) >> 32)) % n_words;
            v[m] = ld<FLAV>(tab + a);
        }
#pragma unroll
        for (int m = 0; m < MLP; m++) acc += v[m];
        s ^= (acc & 1u); // make the next addresses wait for this round's data, like the walk's permeability test
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int FLAV, int MLP>
double run(const uint32_t *tab, uint32_t n_words, uint32_t window, int blocks_per_sm, int sm, int iters, uint32_t *sink)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int grid = blocks_per_sm * sm;
    probe<FLAV, MLP><<<grid, 256>>>(tab, n_words, window, iters / 8, sink); // warm-up
    CK(cudaEventRecord(a));
    probe<FLAV, MLP><<<grid, 256>>>(tab, n_words, window, iters, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    return (double)grid * 256 * iters * MLP / (ms * 1e-3);
}

int main(int argc, char **argv)
{
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, dev));
    const int sm = p.multiProcessorCount;
    int gran = argc > 1 ? atoi(argv[1]) : 0;
    const bool ncu_mode = argc > 2;
    if (gran) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran));
    size_t g = 0;
    CK(cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity));
    fprintf(stderr, "# %s, %d SMs, L2 %d MB, L2 fetch granularity %zu B\n", p.name, sm, p.l2CacheSize >> 20, g);
    uint32_t *sink;
    CK(cudaMalloc(&sink, 4));
    if (ncu_mode) { // one short launch per load flavour, for `ncu --metrics lts__t_sectors_srcunit_tex_op_read.sum,dram__bytes_read.sum,...`
        const uint32_t n_words = (uint32_t)(864ull * 1024 * 1024 / 4);
        uint32_t *tab;
        CK(cudaMalloc(&tab, (size_t)n_words * 4));
        CK(cudaMemset(tab, 0, (size_t)n_words * 4));
#define ONE(F) probe<F, 1><<<8 * sm, 256>>>(tab, n_words, n_words, 256, sink); CK(cudaDeviceSynchronize());
        ONE(0) ONE(1) ONE(2) ONE(3) ONE(4) ONE(5) ONE(6) ONE(7) ONE(8) ONE(9)
#undef ONE
        printf("ncu mode: 10 flavours x %d gathers each\n", 8 * sm * 256 * 256);
        return 0;
    }
    printf("table_MB,window_words,flavour,mlp,blocks_per_sm,Ggathers_per_s,sector_GBps,l2_fetch_B\n");
    const size_t sizes_mb[] = {64, 864, 4096};
    const char *names[] = {"nc", "cg", "nc.noalloc", "nc.L1evict_first"};
    for (size_t mb : sizes_mb) {
        const uint32_t n_words = (uint32_t)(mb * 1024 * 1024 / 4);
        uint32_t *tab;
        CK(cudaMalloc(&tab, (size_t)n_words * 4));
        CK(cudaMemset(tab, 0, (size_t)n_words * 4));
        const uint32_t windows[] = {n_words, 1u << 21, 1u << 14};
        for (uint32_t w : windows) {
            if (w != n_words && mb != 864) continue;
            for (int bps : {4, 8}) {
                const int iters = 2048;
#define ROW(F, M) { double r = run<F, M>(tab, n_words, w, bps, sm, iters, sink); \
                    printf("%zu,%u,%s,%d,%d,%.2f,%.1f,%zu\n", mb, w, names[F], M, bps, r / 1e9, r * 32 / 1e9, g); fflush(stdout); }
                ROW(0, 1) ROW(0, 2) ROW(0, 4)
                ROW(1, 1) ROW(1, 2) ROW(1, 4)
                ROW(2, 1) ROW(2, 4)
                ROW(3, 1) ROW(3, 4)
#undef ROW
            }
        }
        CK(cudaFree(tab));
    }
    return 0;
}
