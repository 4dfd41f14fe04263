// spinwalk_b200/csrc/engine.cu — C-ABI (include/spinwalk_engine.h) over the sm_100a walk kernel.
//
// Host-side counterpart of the device branch of sim::monte_carlo::run (src/sim/monte_carlo.cu:247-337)
// and of save()'s D2H copies (:170-176): owns device memory, validates and stages the sequence,
// launches ONE kernel for all scales, and hands results back in the reference's layouts.
// No CPU fallback exists: every entry point needs a CUDA device.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>

#include "walk_fast.cuh"
#include "walk_kernel.cuh"

using namespace swk;

namespace {

// SWK_TRACE=1: host-side phase timings of a run on stderr (diagnostics only)
struct Trace {
    bool on = getenv("SWK_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void mark(const char *what)
    {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[swk] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

} // namespace

struct swk_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t cstream[2] = {nullptr, nullptr}; // pipelined host runs: slices alternate between two compute streams
    cudaStream_t dstream = nullptr;               // ... and their results are downloaded on this one
    std::vector<cudaEvent_t> ev_slice;
    cudaEvent_t evA = nullptr, ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_join = nullptr;
    std::string err;
    int sm_count = 0;
    size_t smem_optin = 0;

    // phantom
    DevBuf mask, fieldmap, packed;
    bool packed_valid = false;
    bool packed_brick = false; // layout of `packed`: 2 x 2 x 4 bricks instead of row-major
    DevBuf slab;            // SWK_RUN_ZSLAB: packed words of one z plane, [nx][ny]
    bool slab_valid = false, raw_slab_valid = false;
    bool last_used_slab = false; // the last run walked the z slab: swk_probe_gather probes that table
    int z_invariant = -1;   // -1 not checked yet, 0 / 1: mask and field map do not depend on z
    uint64_t dims[3] = {0, 0, 0};
    float fov[3] = {0, 0, 0};
    uint32_t mask_substrates = 0;
    bool has_phantom = false;

    // sequence
    swk_params P{};
    BlobLayout L{};
    std::vector<uint8_t> blob_h;
    DevBuf blob;
    float rf_ph0 = 0.f;
    bool has_sequence = false;

    // spins
    DevBuf xyz0, m0, order, inv_order; // inv_order: spin -> thread slot (unpack_rows_kernel)
    DevBuf state_a, state_b, state_vox;                     // re-binning pauses of long runs: per (scale, spin) walker state
    DevBuf raw_slab;                                        // COMPAT mode, z-invariant phantom: (substrate id, FP32 field bits) of one z plane
    DevBuf mstate;                                          // one walk for all scales (walk_fast.cuh MULTI): magnetisation per (scale, thread slot)
    DevBuf sort_keys_in, sort_keys_out, sort_ids, sort_tmp; // kept between runs: re-sorting after every swk_set_spins must not malloc
    bool order_valid = false;
    uint32_t order_slice = 0; // slice length the order was built for (0 = one slice)
    uint32_t spin_first = 0, n_local = 0;
    bool has_m0 = false, has_spins = false;

    size_t mem_pitch = 0x7fffffff; // cudaDeviceProp::memPitch: the largest pitch cudaMemcpy2D accepts (SWK_MEMPITCH overrides, tests)
    // run state / outputs
    DevBuf scales, M1, XYZ1, T, sums, counters;
    DevBuf stage;     // staging rows of the per-spin results (walk_kernel.cuh WalkArgs::stage)
    DevBuf sums_fx;   // fixed-point ensemble sums the kernels accumulate; converted into `sums` after the walk
    DevBuf scale_tab; // per-scale constants of the FAST kernel
    uint32_t n_scales = 0, n_te = 0, last_slices = 1;
    uint64_t host_rows = 0, host_row_first = 0; // swk_set_host_rows: host arrays are [K][host_rows][...], ours start at row host_row_first
    uint64_t trj = 1;
    int out_flags = 0;
    double *last_sums = nullptr; // device pointer actually used by the last run
    swk_stats stats{};
};

namespace {

int fail(swk_engine *e, int code, const std::string &msg)
{
    if (e) e->err = msg;
    else g_create_error = msg;
    return code;
}

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(e, SWK_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));           \
    } while (0)

void release(DevBuf &b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
}

// (re)allocate exactly `bytes` (0 => free)
int ensure(swk_engine *e, DevBuf &b, size_t bytes)
{
    if (b.bytes == bytes && (b.p || bytes == 0)) return SWK_OK;
    release(b);
    if (bytes == 0) return SWK_OK;
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    if (bytes > free_b) { // ≙ check_memory_size (device_helper.cu:78-101)
        char msg[160];
        snprintf(msg, sizeof msg, "not enough device memory: need %.1f MB, free %.1f MB", bytes / 1048576.0, free_b / 1048576.0);
        return fail(e, SWK_ERR_MEMORY, msg);
    }
    CK(cudaMalloc(&b.p, bytes));
    b.bytes = bytes;
    return SWK_OK;
}

__global__ void mask_max_kernel(const uint8_t *mask, size_t n, unsigned int *out)
{
    unsigned int m = 0;
    const size_t n16 = n / 16;
    const uint4 *v = reinterpret_cast<const uint4 *>(mask);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        uint4 w = __ldg(v + i);
        unsigned int a = __vmaxu4(__vmaxu4(w.x, w.y), __vmaxu4(w.z, w.w));
        a = max(max(a & 0xffu, (a >> 8) & 0xffu), max((a >> 16) & 0xffu, a >> 24));
        m = max(m, a);
    }
    for (size_t i = n16 * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        m = max(m, (unsigned int)mask[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// ---- locality order of the spins -------------------------------------------------------------------------------
// key = (255 - substrate of the start voxel) << 48 | 48-bit Morton code of the start voxel.  Warps then hold spins of one
// substrate that start next to each other: (1) redraw loops behind impermeable walls (kernels.cu:154-160) no longer
// idle the other lanes of a warp, (2) the mask / field sectors a warp gathers are shared in L1/L2.  The start voxel does
// not depend on the FoV scale (positions and FoV scale together, monte_carlo.cu:278-280), so one order serves all scales.
__device__ __forceinline__ uint64_t spread3(uint32_t v)
{ // 16 bits -> every third bit
    uint64_t x = v & 0xffffu;
    x = (x | (x << 32)) & 0x00ff00000000ffffull;
    x = (x | (x << 16)) & 0x00ff0000ff0000ffull;
    x = (x | (x << 8)) & 0xf00f00f00f00f00full;
    x = (x | (x << 4)) & 0x30c30c30c30c30c3ull;
    x = (x | (x << 2)) & 0x9249249249249249ull;
    return x;
}
// Compact voxel word of SWK_MODE_FAST: the FP32 field (Tesla) moved to the NEAREST value whose 4 low mantissa bits spell the substrate
// id (|error| <= 8 ulp = 2^-20 relative).  One 4-byte gather per voxel change instead of a byte + a float from two arrays, and the word
// is used AS the field value (no masking in the walk).
__device__ __forceinline__ uint32_t pack_word(float field, uint32_t ts)
{
    const uint32_t b = __float_as_uint(field), sign = b & 0x80000000u;
    uint32_t mag = b & 0x7fffffffu;
    if ((mag & 0x7f800000u) == 0x7f800000u) return sign | (mag & ~15u) | ts; // inf / nan: keep the class
    const int d = (int)(mag & 15u) - (int)ts;
    mag = (mag & ~15u) | ts;
    if (d > 8) mag += 16u;                    // (a carry into the exponent is still the right value)
    else if (d < -8 && mag >= 16u) mag -= 16u;
    return sign | mag;
}
__global__ void pack_voxels_kernel(const uint8_t *mask, const float *field, size_t n, uint32_t *out)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = pack_word(field[i], mask[i]);
}
// ... in bricks of 2 x 2 x 4 voxels (16 words = one 64-byte fetch unit; walk_fast.cuh table_index); padding words of odd sizes are never read
__global__ void pack_bricks_kernel(const uint8_t *mask, const float *field, uint32_t nx, uint32_t ny, uint32_t nz, uint32_t *out)
{
    const size_t n = (size_t)nx * ny * nz;
    const uint32_t bry = (ny + 1u) >> 1, brz = (nz + 3u) >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t z = (uint32_t)(i % nz), y = (uint32_t)((i / nz) % ny), x = (uint32_t)(i / ((size_t)nz * ny));
        out[(((((size_t)(x >> 1) * bry + (y >> 1)) * brz + (z >> 2)) << 4) | ((x & 1u) << 3) | ((y & 1u) << 2) | (z & 3u))] = pack_word(field[i], mask[i]);
    }
}

// z-slab table: does any voxel differ from the z = 0 voxel of its column?  (one streaming pass over mask and field map)
__global__ void zinv_check_kernel(const uint8_t *mask, const float *field, size_t n, uint32_t nz, unsigned int *differs)
{
    bool d = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t c = i - i % nz;
        d |= (mask[i] != mask[c]) | (__float_as_uint(field[i]) != __float_as_uint(field[c]));
    }
    if (__any_sync(0xffffffffu, d) && (threadIdx.x & 31) == 0) atomicOr(differs, 1u);
}
// ... and the packed words (pack_voxels_kernel) of the z = 0 plane
__global__ void pack_slab_kernel(const uint8_t *mask, const float *field, size_t nxy, uint32_t nz, uint32_t *out)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nxy; i += (size_t)gridDim.x * blockDim.x) out[i] = pack_word(field[i * nz], mask[i * nz]);
}
// ... and, for SWK_MODE_COMPAT, the UNROUNDED pair (substrate id, FP32 field bits) of the z = 0 plane: one 8-byte gather per voxel change
__global__ void raw_slab_kernel(const uint8_t *mask, const float *field, size_t nxy, uint32_t nz, uint2 *out)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nxy; i += (size_t)gridDim.x * blockDim.x)
        out[i] = make_uint2(mask[i * nz], __float_as_uint(field[i * nz]));
}

// With slice_len != 0 the slice number of the spin (id / slice_len) leads the key, so that the sorted order is slice-major
// and a pipelined run can launch (and download) one contiguous id range after the other.
__global__ void sort_keys_kernel(const float *xyz0, uint32_t n, const uint8_t *mask, uint32_t nx, uint32_t ny, uint32_t nz, float ihx,
                                 float ihy, float ihz, uint32_t slice_len, uint64_t *keys, uint32_t *ids)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int vx = max(0, min((int)floorf(xyz0[3 * (size_t)j + 0] * ihx), (int)nx - 1));
    const int vy = max(0, min((int)floorf(xyz0[3 * (size_t)j + 1] * ihy), (int)ny - 1));
    const int vz = max(0, min((int)floorf(xyz0[3 * (size_t)j + 2] * ihz), (int)nz - 1));
    const uint32_t ts = mask[((size_t)vx * ny + vy) * nz + vz];
    const uint64_t slice = slice_len ? j / slice_len : 0u;
    keys[j] = (slice << 56) | ((uint64_t)(255u - ts) << 48) | (spread3(vx) << 2) | (spread3(vy) << 1) | spread3(vz);
    ids[j] = j;
}

// swk_probe_gather: dependent random gathers, no other work (same load instruction as the FAST walk).
__device__ __forceinline__ uint32_t probe_hash(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__global__ void __launch_bounds__(256) gather_probe_kernel(const uint32_t *tab, uint32_t n_words, uint32_t iters, uint32_t *sink)
{
    uint32_t s = probe_hash((blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u), acc = 0;
    for (uint32_t it = 0; it < iters; it++) {
        s = probe_hash(s + 0x9e3779b9u);
        const uint32_t v = ldg_voxel(tab + (uint32_t)(((uint64_t)s * n_words) >> 32));
        acc += v;
        s ^= (v & 1u); // the next address waits for this gather, like the walk's permeability test
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

// Re-binning keys of a paused long run: like sort_keys_kernel, but from the walkers' CURRENT voxel (state_vox) and substrate,
// one segment per scale (the scale index leads the key) when the walks differ between scales.
__global__ void rebin_keys_kernel(const uint32_t *state_vox, const uint4 *state_b, uint32_t n_local, uint32_t n_seg, uint32_t ny, uint32_t nz,
                                  uint64_t *keys, uint32_t *ids)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n_seg * n_local) return;
    const uint32_t v = state_vox[i], ts = state_b[i].w & 0xffu;
    const uint32_t vz = v % nz, vy = (v / nz) % ny, vx = v / (nz * ny);
    keys[i] = ((uint64_t)(i / n_local) << 56) | ((uint64_t)(255u - ts) << 48) | (spread3(vx) << 2) | (spread3(vy) << 1) | spread3(vz);
    ids[i] = (uint32_t)(i % n_local);
}

// Staging rows -> the reference's output layouts (monte_carlo.cu:61-70), rows [r0, r1) of every scale: a streaming pass over the OUTPUT, one
// 16-byte slot per thread; consecutive threads write consecutive elements of M1 / T / XYZ1.  The walkers wrote their rows in thread-slot order
// (whole 512-byte lines per warp); `inv` (inverse of the locality order: spin -> thread slot) un-permutes them here, with 16-byte gathers.
__global__ void unpack_rows_kernel(const uint4 *stage, uint32_t row_slots, uint32_t n_te, size_t S, size_t chunks, size_t r0, size_t r1, uint32_t K,
                                   const uint32_t *inv, float *M1, uint8_t *T, float *XYZ1)
{
    const size_t per_scale = (r1 - r0) * row_slots, total = per_scale * K;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t k = i / per_scale, rem = i - k * per_scale;
        const size_t r = r0 + rem / row_slots, row = k * S + r;
        const uint32_t e = (uint32_t)(rem % row_slots);
        const size_t slot = inv ? (size_t)__ldg(inv + r) : r;
        uint4 v; // one 16-byte gather per output slot; L2::64B: do not pull the whole 128-byte line of somebody else's rows from HBM
        asm("ld.global.nc.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(stage + ((k * chunks + (slot >> 5)) * row_slots + e) * 32u + (slot & 31u)));
        if (e < n_te) {
            if (M1) {
                float *d = M1 + (row * n_te + e) * 3;
                __stcs(d + 0, __uint_as_float(v.x)); __stcs(d + 1, __uint_as_float(v.y)); __stcs(d + 2, __uint_as_float(v.z));
            }
            if (T) T[row * n_te + e] = (uint8_t)v.w;
        } else if (XYZ1) {
            float *d = XYZ1 + row * 3;
            __stcs(d + 0, __uint_as_float(v.x)); __stcs(d + 1, __uint_as_float(v.y)); __stcs(d + 2, __uint_as_float(v.z));
        }
    }
}

// inverse of the locality order: inv[order[j]] = j
__global__ void invert_order_kernel(const uint32_t *order, uint32_t n, uint32_t *inv)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) inv[order[j]] = j;
}

// fixed-point ensemble sums (walk_kernel.cuh echo_sums_add) -> double [K][E][n_sub][4]
__global__ void sums_to_double_kernel(const unsigned long long *fx, size_t n, double *out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long v = (long long)fx[i];
    out[i] = (i & 3u) == 3u ? (double)v : (double)v * (1.0 / (double)kSumScale);
}

// key of the permeability stream (walk_fast.cuh philox2x32_10): a 32-bit fold of the run's seed
uint32_t perm_stream_key(uint64_t seed) { return (uint32_t)seed + 0x9E3779B9u * (uint32_t)(seed >> 32) + 0x7F4A7C15u; }

// swk_debug_rng: the random-number building blocks of the FAST walk, run on caller-supplied inputs (tests/test_rng_gpu.py)
__global__ void debug_rng_kernel(int which, const uint32_t *in, uint32_t n, uint32_t one_bits, uint32_t *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c0 = in[4 * (size_t)i], c1 = in[4 * (size_t)i + 1], c2 = in[4 * (size_t)i + 2], c3 = in[4 * (size_t)i + 3];
    uint32_t o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (which == 0) {
        const uint4 r = philox_fixed(c0, c1, c2, c3);
        o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
    } else if (which == 1) {
        uint32_t key[10];
        for (uint32_t r = 0; r < 10; r++) key[r] = c2 + r * 0x9E3779B9u;
        o[0] = philox2x32_10(c0, c1, key);
        o[1] = __float_as_uint(u01_open1(o[0]));
    } else {
        float a0, a1, a2, b0, b1, b2;
        normals6_fast(make_uint4(c0, c1, c2, c3), one_bits, a0, a1, a2, b0, b1, b2);
        o[0] = __float_as_uint(a0); o[1] = __float_as_uint(a1); o[2] = __float_as_uint(a2);
        o[3] = __float_as_uint(b0); o[4] = __float_as_uint(b1); o[5] = __float_as_uint(b2);
    }
    for (int k = 0; k < 8; k++) out[8 * (size_t)i + k] = o[k];
}

bool ascending(const int32_t *t, uint32_t n)
{
    for (uint32_t i = 1; i < n; i++)
        if (t[i] <= t[i - 1]) return false;
    return true;
}

uint32_t put(std::vector<uint8_t> &b, const void *src, size_t bytes)
{
    while (b.size() % 8) b.push_back(0);
    uint32_t off = (uint32_t)b.size();
    const uint8_t *s = static_cast<const uint8_t *>(src);
    b.insert(b.end(), s, s + bytes);
    return off;
}

} // namespace

extern "C" {

int swk_version(void) { return SWK_VERSION_MAJOR * 100 + SWK_VERSION_MINOR; }

int swk_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int swk_device_info(char *buf, size_t n)
{
    if (!buf || n == 0) return SWK_ERR_INVALID;
    auto version = [](int v) { return std::to_string(v / 1000) + "." + std::to_string((v % 100) / 10); }; // device_helper.cu:41-45
    std::string out;
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess) out = std::string("Error: ") + cudaGetErrorString(err) + "\n";
    else {
        int rt = 0, drv = 0, dev = 0;
        cudaRuntimeGetVersion(&rt);
        cudaDriverGetVersion(&drv);
        out += "The latest version of CUDA supported by the driver: " + version(drv) + ", current CUDA version: " + version(rt) + "\n";
        out += "Number of devices: " + std::to_string(count) + "\n";
        cudaDeviceProp prop{};
        size_t free_b = 0, total_b = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            out += std::string(prop.name) + "\n";
            out += "-Compute Capability: " + std::to_string(prop.major) + "." + std::to_string(prop.minor) + "\n";
            out += "-Free GPU Memory: " + std::to_string(free_b >> 20) + " MB (out of " + std::to_string(total_b >> 20) + " MB)\n";
        }
    }
    snprintf(buf, n, "%s", out.c_str());
    return err == cudaSuccess ? SWK_OK : SWK_ERR_CUDA;
}

const char *swk_last_error(const swk_engine *e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int swk_create(int device_id, swk_engine **out)
{
    swk_engine *e = nullptr;
    if (!out) return fail(nullptr, SWK_ERR_INVALID, "swk_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t ce = cudaGetDeviceCount(&n);
    if (ce != cudaSuccess || n <= 0)
        return fail(nullptr, SWK_ERR_CUDA, std::string("no CUDA device: ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0") +
                                               " (this engine has no CPU fallback)");
    if (device_id < 0 || device_id >= n) {
        char msg[128];
        snprintf(msg, sizeof msg, "device id %d is not available; number of GPUs is %d", device_id, n);
        return fail(nullptr, SWK_ERR_INVALID, msg);
    }
    e = new swk_engine();
    e->device = device_id;
    cudaDeviceProp prop{};
    if ((ce = cudaSetDevice(device_id)) != cudaSuccess || (ce = cudaGetDeviceProperties(&prop, device_id)) != cudaSuccess ||
        (ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (ce = cudaEventCreate(&e->evA)) != cudaSuccess || (ce = cudaEventCreate(&e->ev0)) != cudaSuccess || (ce = cudaEventCreate(&e->ev1)) != cudaSuccess) {
        std::string m = std::string("swk_create: ") + cudaGetErrorString(ce);
        delete e;
        return fail(nullptr, SWK_ERR_CUDA, m);
    }
    e->sm_count = prop.multiProcessorCount;
    e->smem_optin = prop.sharedMemPerBlockOptin;
    e->mem_pitch = prop.memPitch;
    if (const char *ev = getenv("SWK_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(ev)); // experiment (profiles/README.md): 32 / 64 / 128
    if (const char *ev = getenv("SWK_MEMPITCH")) e->mem_pitch = (size_t)strtoull(ev, nullptr, 10); // test hook: exercise the per-scale copies
    *out = e;
    return SWK_OK;
}

void swk_destroy(swk_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    for (DevBuf *b : {&e->mask, &e->fieldmap, &e->packed, &e->blob, &e->xyz0, &e->m0, &e->order, &e->scales, &e->M1, &e->XYZ1, &e->T, &e->sums, &e->counters,
                      &e->sort_keys_in, &e->sort_keys_out, &e->sort_ids, &e->sort_tmp, &e->state_a, &e->state_b, &e->state_vox, &e->mstate, &e->raw_slab, &e->slab, &e->stage, &e->sums_fx, &e->scale_tab, &e->inv_order})
        release(*b);
    if (e->evA) cudaEventDestroy(e->evA);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    for (cudaEvent_t ev : e->ev_slice) cudaEventDestroy(ev);
    for (cudaStream_t st : {e->cstream[0], e->cstream[1], e->dstream})
        if (st) cudaStreamDestroy(st);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int swk_prepare(swk_params *p, float RF_FA0_deg, float T1_0_ms, const double *diffusivity_m2s, uint32_t n, double *sigma_out)
{
    if (!p || p->timestep_us <= 0) return SWK_ERR_INVALID;
    p->c = cosf(RF_FA0_deg * kDeg2Rad); // simulation_parameters.cuh:229-230
    p->s = sinf(RF_FA0_deg * kDeg2Rad);
    p->n_timepoints = (uint32_t)(p->TR_us / p->timestep_us);
    for (uint32_t i = 0; i < n && diffusivity_m2s && sigma_out; i++)
        sigma_out[i] = 1e-3 * sqrt(2. * diffusivity_m2s[i] * p->timestep_us); // :236-237
    if (p->n_dummy_scan < 0) p->n_dummy_scan = 5.0 * T1_0_ms / float(p->TR_us * 1e-3); // :239-242
    return SWK_OK;
}

int swk_set_phantom(swk_engine *e, const uint8_t *mask, const float *fieldmap_T, const uint64_t dims[3], const float fov_m[3], int on_device)
{
    if (!e) return SWK_ERR_INVALID;
    if (!mask || !dims || !fov_m) return fail(e, SWK_ERR_INVALID, "swk_set_phantom: mask, dims and fov are mandatory");
    for (int i = 0; i < 3; i++)
        if (dims[i] == 0 || dims[i] > 0x7fffffffull || !(fov_m[i] > 0.f)) return fail(e, SWK_ERR_INVALID, "swk_set_phantom: bad dims or fov");
    CK(cudaSetDevice(e->device));
    const size_t V = (size_t)dims[0] * dims[1] * dims[2];
    e->has_phantom = false;
    e->order_valid = false;
    release(e->mask); // ≙ cleanup_device (monte_carlo.cu:86-95)
    release(e->fieldmap);
    release(e->packed);
    e->packed_valid = false;
    release(e->slab);
    e->slab_valid = false; e->raw_slab_valid = false;
    e->z_invariant = -1;
    int rc;
    if ((rc = ensure(e, e->mask, V)) != SWK_OK) return rc;
    if (fieldmap_T && (rc = ensure(e, e->fieldmap, V * sizeof(float))) != SWK_OK) return rc;
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    CK(cudaMemcpyAsync(e->mask.p, mask, V, kind, e->stream));
    if (fieldmap_T) CK(cudaMemcpyAsync(e->fieldmap.p, fieldmap_T, V * sizeof(float), kind, e->stream));
    // number of substrates present = max(mask)+1 (monte_carlo.cu:113)
    if ((rc = ensure(e, e->counters, 8 * sizeof(unsigned long long))) != SWK_OK) return rc;
    CK(cudaMemsetAsync(e->counters.p, 0, e->counters.bytes, e->stream));
    mask_max_kernel<<<e->sm_count * 8, 256, 0, e->stream>>>(static_cast<const uint8_t *>(e->mask.p), V, static_cast<unsigned int *>(e->counters.p));
    CK(cudaGetLastError());
    unsigned int mx = 0;
    CK(cudaMemcpyAsync(&mx, e->counters.p, sizeof mx, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->mask_substrates = mx + 1;
    for (int i = 0; i < 3; i++) { e->dims[i] = dims[i]; e->fov[i] = fov_m[i]; }
    e->has_phantom = true;
    return SWK_OK;
}

int swk_set_sequence(swk_engine *e, const swk_params *p, const swk_tables *t)
{
    if (!e) return SWK_ERR_INVALID;
    if (!p || !t) return fail(e, SWK_ERR_INVALID, "swk_set_sequence: NULL argument");
    // ---- validation ≙ config_reader::check (config_reader.cpp:195-324) ----
    const uint32_t ns = p->n_substrate;
    if (ns < 1 || ns > 255) return fail(e, SWK_ERR_INVALID, "at least one (and at most 255) substrates are required");
    if (t->n_step_sigma != ns || t->n_T1 != ns || t->n_T2 != ns || t->n_pXY != ns * ns || !t->step_sigma_m || !t->T1_ms || !t->T2_ms || !t->pXY)
        return fail(e, SWK_ERR_INVALID, "DIFFUSIVITY, T1, T2 must have n_substrate entries and P_XY n_substrate^2");
    if (t->n_RF < 1 || t->n_RF_FA != t->n_RF || t->n_RF_PH != t->n_RF || !t->RF_tp || !t->RF_FA_deg || !t->RF_PH_deg)
        return fail(e, SWK_ERR_INVALID, "RF_FA, RF_PH and RF_T must have the same number of elements (>= 1)");
    if (t->RF_tp[0] != 0) return fail(e, SWK_ERR_INVALID, "the first RF start time must be 0");
    if (t->n_dephasing_deg != t->n_dephasing) return fail(e, SWK_ERR_INVALID, "DEPHASING and DEPHASING_T must have the same number of elements");
    if (t->n_gradX != t->n_gradient || t->n_gradY != t->n_gradient || t->n_gradZ != t->n_gradient)
        return fail(e, SWK_ERR_INVALID, "GRADIENT_X/Y/Z and GRADIENT_T must have the same number of elements");
    if (t->n_RF > 65535 || t->n_TE > 65535 || t->n_dephasing > 65535 || t->n_gradient > 65535)
        return fail(e, SWK_ERR_INVALID, "event tables are limited to 65535 entries (uint16 counters, kernels.cu:125)");
    if (!ascending(t->RF_tp, t->n_RF) || !ascending(t->TE_tp, t->n_TE) || !ascending(t->dephasing_tp, t->n_dephasing) ||
        !ascending(t->gradient_tp, t->n_gradient))
        return fail(e, SWK_ERR_INVALID, "RF_T, TE, DEPHASING_T and GRADIENT_T must be sorted in strictly ascending order");
    for (uint32_t i = 0; i < t->n_TE; i++)
        if (t->TE_tp[i] < 0) return fail(e, SWK_ERR_INVALID, "TE must be >= 0");
    for (uint32_t i = 0; i < t->n_dephasing; i++)
        if (t->dephasing_tp[i] < 0) return fail(e, SWK_ERR_INVALID, "DEPHASING_T must be >= 0");
    for (uint32_t i = 0; i < t->n_gradient; i++)
        if (t->gradient_tp[i] < 0) return fail(e, SWK_ERR_INVALID, "GRADIENT_T must be >= 0");
    if (p->timestep_us <= 0 || p->TR_us <= 0 || p->n_timepoints == 0) return fail(e, SWK_ERR_INVALID, "TR and TIME_STEP must be positive");
    if (p->n_dummy_scan < 0) return fail(e, SWK_ERR_INVALID, "n_dummy_scan must be resolved (>= 0): call swk_prepare");
    if (p->seed == 0) return fail(e, SWK_ERR_INVALID, "seed must be non-zero (the caller resolves SEED = 0 like parameters::prepare)");
    if (p->n_spins == 0) return fail(e, SWK_ERR_INVALID, "NUMBER_OF_SPINS must be positive");
    CK(cudaSetDevice(e->device));

    // ---- merged event timeline ----
    struct Ev { int32_t time; uint32_t mask; };
    std::vector<Ev> evs;
    auto add = [&](const int32_t *tp, uint32_t n, uint32_t first, uint32_t bit) {
        for (uint32_t i = first; i < n; i++) evs.push_back({tp[i], bit});
    };
    add(t->dephasing_tp, t->n_dephasing, 0, EV_DEPH);
    add(t->gradient_tp, t->n_gradient, 0, EV_GRAD);
    add(t->RF_tp, t->n_RF, 1, EV_RF); // pulse 0 is applied at the start of every TR (kernels.cu:117,125)
    add(t->TE_tp, t->n_TE, 0, EV_ECHO);
    std::stable_sort(evs.begin(), evs.end(), [](const Ev &a, const Ev &b) { return a.time < b.time; });
    std::vector<int32_t> tl_time;
    std::vector<uint32_t> tl_mask;
    for (const Ev &v : evs) {
        if (!tl_time.empty() && tl_time.back() == v.time) tl_mask.back() |= v.mask;
        else { tl_time.push_back(v.time); tl_mask.push_back(v.mask); }
    }

    // ---- tables: double-precision sin/cos of the flip angles (kernels.cuh:203-206), T1/T2 in seconds ----
    std::vector<float> rf_s(t->n_RF), rf_c(t->n_RF), T1s(ns), T2s(ns);
    for (uint32_t i = 0; i < t->n_RF; i++) {
        rf_s[i] = (float)sin(t->RF_FA_deg[i] * kDeg2Rad);
        rf_c[i] = (float)cos(t->RF_FA_deg[i] * kDeg2Rad);
    }
    for (uint32_t i = 0; i < ns; i++) {
        T1s[i] = (float)(t->T1_ms[i] * 1e-3); // kernels.cu:167-168
        T2s[i] = (float)(t->T2_ms[i] * 1e-3);
    }
    std::vector<uint8_t> b;
    BlobLayout L{};
    const float zero = 0.f;
    L.n_tl = (uint32_t)tl_time.size();
    L.tl_time = put(b, tl_time.empty() ? (const void *)&zero : tl_time.data(), tl_time.size() * 4);
    L.tl_mask = put(b, tl_mask.empty() ? (const void *)&zero : tl_mask.data(), tl_mask.size() * 4);
    std::vector<uint32_t> tl_run(tl_time.size(), 0u); // gradient-only entries at consecutive timepoints (e.g. a PGSE lobe)
    for (size_t i = tl_time.size(); i-- > 0;)
        if (tl_mask[i] == EV_GRAD)
            tl_run[i] = (i + 1 < tl_time.size() && tl_mask[i + 1] == EV_GRAD && tl_time[i + 1] == tl_time[i] + 1) ? tl_run[i + 1] + 1u : 1u;
    L.tl_run = put(b, tl_run.empty() ? (const void *)&zero : tl_run.data(), tl_run.size() * 4);
    L.n_rf = t->n_RF;
    L.rf_s = put(b, rf_s.data(), rf_s.size() * 4);
    L.rf_c = put(b, rf_c.data(), rf_c.size() * 4);
    L.rf_ph = put(b, t->RF_PH_deg, t->n_RF * 4);
    L.n_deph = t->n_dephasing;
    L.deph_deg = put(b, t->dephasing_deg ? (const void *)t->dephasing_deg : &zero, t->n_dephasing * 4);
    L.n_grad = t->n_gradient;
    L.gx = put(b, t->gradX_mTm ? (const void *)t->gradX_mTm : &zero, t->n_gradient * 4);
    L.gy = put(b, t->gradY_mTm ? (const void *)t->gradY_mTm : &zero, t->n_gradient * 4);
    L.gz = put(b, t->gradZ_mTm ? (const void *)t->gradZ_mTm : &zero, t->n_gradient * 4);
    L.n_sub = ns;
    L.sigma = put(b, t->step_sigma_m, ns * 8);
    L.T1s = put(b, T1s.data(), ns * 4);
    L.T2s = put(b, T2s.data(), ns * 4);
    L.pXY = put(b, t->pXY, ns * ns * 4);
    while (b.size() % 16) b.push_back(0);
    L.bytes = (uint32_t)b.size();

    int rc;
    if ((rc = ensure(e, e->blob, b.size())) != SWK_OK) return rc;
    CK(cudaMemcpyAsync(e->blob.p, b.data(), b.size(), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->blob_h.swap(b);
    e->L = L;
    e->P = *p;
    e->n_te = t->n_TE;
    e->rf_ph0 = t->RF_PH_deg[0];
    e->trj = p->record_trajectory ? (uint64_t)p->n_timepoints * (uint64_t)(p->n_dummy_scan + 1) : 1;
    e->has_sequence = true;
    return SWK_OK;
}

int swk_set_spins(swk_engine *e, const float *XYZ0, const float *M0, uint32_t spin_first, uint32_t n_local)
{
    if (!e) return SWK_ERR_INVALID;
    if (n_local == 0) return fail(e, SWK_ERR_INVALID, "swk_set_spins: n_local must be positive");
    if (!XYZ0 && (!e->has_phantom || !e->has_sequence))
        return fail(e, SWK_ERR_STATE, "swk_set_spins: device-side positions need the phantom (FoV) and the sequence (seed) first");
    CK(cudaSetDevice(e->device));
    int rc;
    const size_t bytes = (size_t)n_local * 3 * sizeof(float);
    if ((rc = ensure(e, e->xyz0, bytes)) != SWK_OK) return rc;
    if (XYZ0) {
        CK(cudaMemcpyAsync(e->xyz0.p, XYZ0, bytes, cudaMemcpyHostToDevice, e->stream));
    } else {
        init_positions_kernel<<<(n_local + 255) / 256, 256, 0, e->stream>>>(static_cast<float *>(e->xyz0.p), n_local, spin_first, e->P.seed,
                                                                           e->fov[0], e->fov[1], e->fov[2]);
        CK(cudaGetLastError());
    }
    e->has_m0 = (M0 != nullptr);
    if (M0) {
        if ((rc = ensure(e, e->m0, bytes)) != SWK_OK) return rc;
        CK(cudaMemcpyAsync(e->m0.p, M0, bytes, cudaMemcpyHostToDevice, e->stream));
    } else {
        release(e->m0);
    }
    CK(cudaStreamSynchronize(e->stream)); // the caller may reuse its host buffers
    e->spin_first = spin_first;
    e->n_local = n_local;
    e->has_spins = true;
    e->order_valid = false;
    return SWK_OK;
}

} // extern "C"

// Host destinations of a pipelined run (swk_run): results of slice i are copied back while slice i+1 computes.
struct HostOut {
    float *M1 = nullptr;
    float *XYZ1 = nullptr;
    uint8_t *T = nullptr;
};

// Rows [.., +width) of K scale blocks, device -> host.  cudaMemcpy2DAsync rejects pitches above cudaDeviceProp::memPitch (2^31 - 1):
// a per-scale block of M1 or XYZ1 of 2 GiB or more (1e8 spins x 2 echoes; trajectories) would fail with "invalid pitch" after the whole
// simulation has run.  One plain copy when the blocks are contiguous on both sides, one 2-D copy when the pitches fit, else one plain
// copy per scale.
static cudaError_t copy_rows(char *dst, size_t dpitch, const char *src, size_t spitch, size_t width, size_t K, size_t pitch_limit, cudaStream_t st)
{
    if (width == 0 || K == 0) return cudaSuccess;
    if (K == 1 || (dpitch == width && spitch == width)) return cudaMemcpyAsync(dst, src, width * K, cudaMemcpyDeviceToHost, st);
    if (dpitch <= pitch_limit && spitch <= pitch_limit) return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, K, cudaMemcpyDeviceToHost, st);
    for (size_t k = 0; k < K; k++) {
        const cudaError_t ce = cudaMemcpyAsync(dst + k * dpitch, src + k * spitch, width, cudaMemcpyDeviceToHost, st);
        if (ce != cudaSuccess) return ce;
    }
    return cudaSuccess;
}

// Per-scale constants of the FAST kernel (walk_fast.cuh ScaleConst): fraction bits of the fixed-point position, step sigma per
// (substrate, axis) in fixed-point units, unit conversions.  Double precision, once per run.
static void scale_constants(const swk_engine *e, const float *scales, uint32_t K, int scale_type, std::vector<uint8_t> &tab, uint32_t &stride)
{
    const uint32_t ns = e->P.n_substrate;
    const uint32_t sgt_off = (uint32_t)((sizeof(ScaleConst) + 15) / 16 * 16);
    stride = (uint32_t)((sgt_off + 3 * ns * sizeof(float) + 15) / 16 * 16);
    tab.assign((size_t)stride * K, 0);
    const double *tsig = reinterpret_cast<const double *>(e->blob_h.data() + e->L.sigma);
    const uint32_t n3[3] = {(uint32_t)e->dims[0], (uint32_t)e->dims[1], (uint32_t)e->dims[2]};
    double inv_h[3]; // grid units per metre at scale 1
    for (int i = 0; i < 3; i++) inv_h[i] = (double)n3[i] / (double)e->fov[i];
    for (uint32_t k = 0; k < K; k++) {
        ScaleConst sc{};
        sc.fscale = 1.f; sc.gscale = 1.f; sc.lin_pc = e->P.linear_phase_cycling;
        if (scale_type == SWK_SCALE_FOV) sc.fscale = scales[k];
        else if (scale_type == SWK_SCALE_GRADIENT) sc.gscale = scales[k];
        else sc.lin_pc = e->P.linear_phase_cycling * scales[k]; // monte_carlo.cu:303 (one FP32 product)
        double smax = 0.;
        for (uint32_t s = 0; s < ns; s++)
            for (int i = 0; i < 3; i++) smax = std::max(smax, tsig[s] * inv_h[i] / (double)sc.fscale);
        const uint32_t nmax = std::max(n3[0], std::max(n3[1], n3[2]));
        int f = 22;
        while (f > 0 && ((double)(nmax + 1u) * (double)(1u << f) + 4194304. > 4294967296.)) f--; // (b) of walk_fast.cuh
        while (f > 0 && 5.7 * smax * (double)(1u << f) >= 4194304.) f--;                           // (a)
        sc.fb = (uint32_t)f;
        sc.sgt_off = sgt_off;
        const double two_f = (double)(1u << f);
        for (int i = 0; i < 3; i++) {
            sc.unit_m[i] = (double)sc.fscale / (inv_h[i] * two_f);
            sc.pos_k[i] = inv_h[i] * two_f;
            sc.pos_hi[i] = (double)n3[i] * two_f - 1.;
            sc.umk[i] = (float)((double)sc.fscale / (inv_h[i] * two_f) * 1e-3 * (double)e->P.timestep_us * 1e-6 * kGamma * kRad2Deg); // kernels.cu:185
        }
        uint8_t *rec = tab.data() + (size_t)k * stride;
        memcpy(rec, &sc, sizeof sc);
        float *sgt = reinterpret_cast<float *>(rec + sgt_off);
        for (uint32_t s = 0; s < ns; s++)
            for (int i = 0; i < 3; i++) sgt[3 * s + i] = (float)(tsig[s] * inv_h[i] / (double)sc.fscale * two_f);
    }
}

// scales per block of the SHARED FAST kernel: the group size in [5, 16] that wastes the fewest warp slots in the last group, larger
// groups first (more walkers share one generation of normals); runs with fewer than 5 scales take them all in one group.
static uint32_t shared_group(uint32_t K)
{
    const uint32_t gmax = SWK_FAST_SHARED_MAXT / 32;
    if (K <= gmax) return K;
    uint32_t best = gmax, best_waste = 0xffffffffu;
    for (uint32_t g = gmax; g >= 5; g--) {
        const uint32_t waste = (K + g - 1) / g * g - K;
        if (waste < best_waste) { best = g; best_waste = waste; }
    }
    return best;
}

typedef void (*walk_fn)(const WalkArgs);

template <bool SHARED>
static walk_fn pick_fast(int vox, bool gruns, bool record, bool stats)
{
#define SWK_PICK2(V, G) (record ? (stats ? walk_fast_kernel<true, true, V, G, SHARED> : walk_fast_kernel<false, true, V, G, SHARED>) \
                                : (stats ? walk_fast_kernel<true, false, V, G, SHARED> : walk_fast_kernel<false, false, V, G, SHARED>))
#define SWK_PICK(V) (gruns ? SWK_PICK2(V, true) : SWK_PICK2(V, false))
    return vox == VOX_PACKED ? SWK_PICK(VOX_PACKED) : (vox == VOX_SLAB ? SWK_PICK(VOX_SLAB) : (vox == VOX_SPLIT ? SWK_PICK(VOX_SPLIT) : SWK_PICK(VOX_MASK)));
#undef SWK_PICK
#undef SWK_PICK2
}
// one walk for all gradient / phase-cycling scales (walk_fast.cuh MULTI): PRIVATE geometry, no trajectory recording
static walk_fn pick_fast_multi(int vox, bool gruns, bool stats)
{
#define SWK_PICK2(V, G) (stats ? walk_fast_kernel<true, false, V, G, false, true> : walk_fast_kernel<false, false, V, G, false, true>)
#define SWK_PICK(V) (gruns ? SWK_PICK2(V, true) : SWK_PICK2(V, false))
    return vox == VOX_PACKED ? SWK_PICK(VOX_PACKED) : (vox == VOX_SLAB ? SWK_PICK(VOX_SLAB) : (vox == VOX_SPLIT ? SWK_PICK(VOX_SPLIT) : SWK_PICK(VOX_MASK)));
#undef SWK_PICK
#undef SWK_PICK2
}

static int run_impl(swk_engine *e, const float *scales, uint32_t n_scales, int scale_type, int mode, int flags, double *d_sums,
                    uint32_t n_slices, const HostOut *host)
{
    if (!e) return SWK_ERR_INVALID;
    if (!e->has_phantom) return fail(e, SWK_ERR_STATE, "swk_run_device: no phantom (swk_set_phantom)");
    if (!e->has_sequence) return fail(e, SWK_ERR_STATE, "swk_run_device: no sequence (swk_set_sequence)");
    if (!e->has_spins) return fail(e, SWK_ERR_STATE, "swk_run_device: no spins (swk_set_spins)");
    if (!scales || n_scales == 0) return fail(e, SWK_ERR_INVALID, "swk_run_device: at least one scale is required");
    if (scale_type < SWK_SCALE_FOV || scale_type > SWK_SCALE_PHASE_CYCLING) return fail(e, SWK_ERR_INVALID, "WHAT_TO_SCALE must be 0, 1 or 2");
    if (mode != SWK_MODE_COMPAT && mode != SWK_MODE_FAST) return fail(e, SWK_ERR_INVALID, "unknown mode");
    if (e->mask_substrates > e->P.n_substrate) { // monte_carlo.cu:113-118
        char msg[200];
        snprintf(msg, sizeof msg, "the number of substrate types in the mask does not match the config: %u vs %u", e->mask_substrates,
                 e->P.n_substrate);
        return fail(e, SWK_ERR_SUBSTRATE, msg);
    }
    if ((uint64_t)e->spin_first + e->n_local > e->P.n_spins) return fail(e, SWK_ERR_INVALID, "spin shard exceeds the global number of spins");
    if (e->host_rows && e->host_row_first + e->n_local > e->host_rows) return fail(e, SWK_ERR_INVALID, "swk_set_host_rows: this engine's rows exceed the host arrays");
    for (uint32_t i = 0; i < n_scales; i++)
        if (scale_type == SWK_SCALE_FOV && !(scales[i] > 0.f)) return fail(e, SWK_ERR_INVALID, "FoV scales must be positive");
    CK(cudaSetDevice(e->device));

    const size_t K = n_scales, S = e->n_local, E = e->n_te, ns = e->P.n_substrate;
    const bool record = e->P.record_trajectory != 0;
    // staging rows carry M1, T and the final position; recorded trajectories are written straight into XYZ1
    const bool want_stage = (flags & (SWK_OUT_M1 | SWK_OUT_T)) || ((flags & SWK_OUT_XYZ1) && !record);
    const size_t stage_row = E + 1;
    int rc;
    if ((rc = ensure(e, e->scales, K * sizeof(float))) != SWK_OK) return rc;
    if ((rc = ensure(e, e->M1, (flags & SWK_OUT_M1) ? K * S * E * 3 * sizeof(float) : 0)) != SWK_OK) return rc;
    if ((rc = ensure(e, e->XYZ1, (flags & SWK_OUT_XYZ1) ? K * S * e->trj * 3 * sizeof(float) : 0)) != SWK_OK) return rc;
    if ((rc = ensure(e, e->T, (flags & SWK_OUT_T) ? K * S * E : 0)) != SWK_OK) return rc;
    const size_t stage_chunks = (S + 31) / 32;
    if ((rc = ensure(e, e->stage, want_stage ? K * stage_chunks * stage_row * 32 * sizeof(uint4) : 0)) != SWK_OK) return rc;
    if ((rc = ensure(e, e->sums, std::max<size_t>(K * E * ns * 4, 1) * sizeof(double))) != SWK_OK) return rc;
    if ((rc = ensure(e, e->sums_fx, std::max<size_t>(K * E * ns * 4, 1) * sizeof(unsigned long long))) != SWK_OK) return rc;
    if ((rc = ensure(e, e->counters, 8 * sizeof(unsigned long long))) != SWK_OK) return rc;
    double *sums = d_sums ? d_sums : static_cast<double *>(e->sums.p);

    // ---- slices: contiguous id ranges launched one after the other (1 = the whole shard in one launch) ----
    if (n_slices < 1) n_slices = 1;
    n_slices = std::min<uint32_t>(n_slices, 256u); // the slice number is the top byte of the sort key
    uint32_t slice_len = (uint32_t)S;
    if (n_slices > 1) {
        slice_len = (uint32_t)(((S + n_slices - 1) / n_slices + kBlock - 1) / kBlock * kBlock);
        n_slices = (uint32_t)((S + slice_len - 1) / slice_len);
    }
    if (n_slices > 1) {
        for (int c = 0; c < 2; c++)
            if (!e->cstream[c]) CK(cudaStreamCreateWithFlags(&e->cstream[c], cudaStreamNonBlocking));
        if (!e->dstream) CK(cudaStreamCreateWithFlags(&e->dstream, cudaStreamNonBlocking));
        while (e->ev_slice.size() < n_slices) {
            cudaEvent_t ev;
            CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            e->ev_slice.push_back(ev);
        }
    }

    Trace tr;
    CK(cudaMemcpyAsync(e->scales.p, scales, K * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    CK(cudaEventRecord(e->evA, e->stream));
    // Every element of M1 / T / XYZ1 is written by the unpack pass (abandoned spins and echoes that never fire are staged as zeros, like the
    // reference's zero-initialised outputs, monte_carlo.cu:256,259-260); only recorded trajectories, which walkers write slot by slot, start at zero.
    if (e->XYZ1.p && record) CK(cudaMemsetAsync(e->XYZ1.p, 0, e->XYZ1.bytes, e->stream));
    if (E * ns) CK(cudaMemsetAsync(e->sums_fx.p, 0, K * E * ns * 4 * sizeof(unsigned long long), e->stream));
    CK(cudaMemsetAsync(e->counters.p, 0, e->counters.bytes, e->stream));

    tr.mark("alloc + memset enqueue");
    uint32_t extra_launches = 0;
    const uint32_t want_order_slice = n_slices > 1 ? slice_len : 0u;
    if (flags & SWK_RUN_NO_SORT) {
        release(e->order);
        e->order_valid = false;
    } else if (!e->order_valid || e->order_slice != want_order_slice) {
        DevBuf &keys_in = e->sort_keys_in, &keys_out = e->sort_keys_out, &ids_in = e->sort_ids, &tmp = e->sort_tmp;
        int rs = SWK_OK;
        size_t tmp_bytes = 0;
        if ((rs = ensure(e, e->order, S * sizeof(uint32_t))) == SWK_OK && (rs = ensure(e, keys_in, S * 8)) == SWK_OK &&
            (rs = ensure(e, keys_out, S * 8)) == SWK_OK && (rs = ensure(e, ids_in, S * 4)) == SWK_OK) {
            sort_keys_kernel<<<(unsigned)((S + 255) / 256), 256, 0, e->stream>>>(
                static_cast<const float *>(e->xyz0.p), (uint32_t)S, static_cast<const uint8_t *>(e->mask.p), (uint32_t)e->dims[0],
                (uint32_t)e->dims[1], (uint32_t)e->dims[2], (float)e->dims[0] / e->fov[0], (float)e->dims[1] / e->fov[1],
                (float)e->dims[2] / e->fov[2], want_order_slice, static_cast<uint64_t *>(keys_in.p), static_cast<uint32_t *>(ids_in.p));
            cudaError_t ce = cudaGetLastError();
            if (ce == cudaSuccess)
                ce = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, static_cast<uint64_t *>(keys_in.p), static_cast<uint64_t *>(keys_out.p),
                                                     static_cast<uint32_t *>(ids_in.p), static_cast<uint32_t *>(e->order.p), (int)S, 0, 64, e->stream);
            if (ce == cudaSuccess && (rs = ensure(e, tmp, tmp_bytes)) == SWK_OK)
                ce = cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, static_cast<uint64_t *>(keys_in.p), static_cast<uint64_t *>(keys_out.p),
                                                     static_cast<uint32_t *>(ids_in.p), static_cast<uint32_t *>(e->order.p), (int)S, 0, 64, e->stream);
            if (ce != cudaSuccess) rs = fail(e, SWK_ERR_CUDA, std::string("spin ordering: ") + cudaGetErrorString(ce));
        }
        if (rs != SWK_OK) return rs;
        if ((rs = ensure(e, e->inv_order, S * sizeof(uint32_t))) != SWK_OK) return rs;
        invert_order_kernel<<<(unsigned)((S + 255) / 256), 256, 0, e->stream>>>(static_cast<const uint32_t *>(e->order.p), (uint32_t)S, static_cast<uint32_t *>(e->inv_order.p));
        CK(cudaGetLastError());
        e->order_valid = true;
        e->order_slice = want_order_slice;
        extra_launches += 2;
    }

    tr.mark("spin ordering");
    // compact voxel words for the FAST walk (built once per phantom)
    const bool want_packed = mode == SWK_MODE_FAST && e->fieldmap.p && e->mask_substrates <= 16 && !(flags & SWK_RUN_NO_PACK);
    // a phantom that does not depend on z (every cylinder phantom of `spinwalk phantom -c`) is walked on its [nx][ny] slab: the same
    // words, hence the same results bit for bit, from a table nz times smaller (L1/L2 resident).  Checked once per phantom on the device.
    // COMPAT mode walks the same plane through raw (substrate id, FP32 field) pairs: the very values of the full arrays, so T / XYZ1 / M1 stay
    // bit-identical to the reference's (tests/test_engine_gpu.py), from an L2-resident table and with one gather instead of two.
    bool use_slab = false, use_raw_slab = false;
    const bool compat_slab = mode == SWK_MODE_COMPAT && e->fieldmap.p != nullptr;
    if ((want_packed || compat_slab) && !(flags & SWK_RUN_NO_ZSLAB) && getenv("SWK_NO_ZSLAB") == nullptr && e->dims[2] > 1) {
        const size_t V = (size_t)(e->dims[0] * e->dims[1] * e->dims[2]), nxy = (size_t)(e->dims[0] * e->dims[1]);
        if (e->z_invariant < 0) {
            unsigned int differs = 0;
            unsigned int *flag = reinterpret_cast<unsigned int *>(static_cast<unsigned long long *>(e->counters.p) + 7); // zeroed above, unused by the walk
            zinv_check_kernel<<<e->sm_count * 16, 256, 0, e->stream>>>(static_cast<const uint8_t *>(e->mask.p), static_cast<const float *>(e->fieldmap.p), V,
                                                                      (uint32_t)e->dims[2], flag);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(&differs, flag, sizeof differs, cudaMemcpyDeviceToHost, e->stream));
            CK(cudaStreamSynchronize(e->stream));
            e->z_invariant = differs ? 0 : 1;
            extra_launches++;
        }
        if (e->z_invariant == 1 && compat_slab) {
            if (!e->raw_slab_valid) {
                if ((rc = ensure(e, e->raw_slab, nxy * sizeof(uint2))) != SWK_OK) return rc;
                raw_slab_kernel<<<e->sm_count * 4, 256, 0, e->stream>>>(static_cast<const uint8_t *>(e->mask.p), static_cast<const float *>(e->fieldmap.p), nxy,
                                                                       (uint32_t)e->dims[2], static_cast<uint2 *>(e->raw_slab.p));
                CK(cudaGetLastError());
                e->raw_slab_valid = true;
                extra_launches++;
            }
            use_raw_slab = true;
        } else if (e->z_invariant == 1) {
            if (!e->slab_valid) {
                if ((rc = ensure(e, e->slab, nxy * sizeof(uint32_t))) != SWK_OK) return rc;
                pack_slab_kernel<<<e->sm_count * 4, 256, 0, e->stream>>>(static_cast<const uint8_t *>(e->mask.p), static_cast<const float *>(e->fieldmap.p), nxy,
                                                                        (uint32_t)e->dims[2], static_cast<uint32_t *>(e->slab.p));
                CK(cudaGetLastError());
                e->slab_valid = true;
                extra_launches++;
            }
            use_slab = true;
        }
    }
    const bool want_brick = getenv("SWK_BRICK") != nullptr && atoi(getenv("SWK_BRICK")) != 0; // experiment (profiles/README.md): 2 x 2 x 4 bricks
    if (want_packed && !use_slab && (!e->packed_valid || e->packed_brick != want_brick)) {
        const size_t V = (size_t)(e->dims[0] * e->dims[1] * e->dims[2]);
        const size_t words = want_brick ? (size_t)((e->dims[0] + 1) / 2) * ((e->dims[1] + 1) / 2) * ((e->dims[2] + 3) / 4) * 16 : V;
        if (words >= (1ull << 32)) return fail(e, SWK_ERR_INVALID, "SWK_MODE_FAST indexes voxels with 32 bits: phantom too large");
        if ((rc = ensure(e, e->packed, words * sizeof(uint32_t))) != SWK_OK) return rc;
        if (want_brick)
            pack_bricks_kernel<<<e->sm_count * 16, 256, 0, e->stream>>>(static_cast<const uint8_t *>(e->mask.p), static_cast<const float *>(e->fieldmap.p), (uint32_t)e->dims[0],
                                                                      (uint32_t)e->dims[1], (uint32_t)e->dims[2], static_cast<uint32_t *>(e->packed.p));
        else
            pack_voxels_kernel<<<e->sm_count * 16, 256, 0, e->stream>>>(static_cast<const uint8_t *>(e->mask.p), static_cast<const float *>(e->fieldmap.p), V,
                                                                      static_cast<uint32_t *>(e->packed.p));
        CK(cudaGetLastError());
        e->packed_valid = true;
        e->packed_brick = want_brick;
        extra_launches++;
    }

    WalkArgs A{};
    A.mask = static_cast<const uint8_t *>(e->mask.p);
    A.fieldmap = static_cast<const float *>(e->fieldmap.p);
    A.packed = want_packed ? static_cast<const uint32_t *>(use_slab ? e->slab.p : e->packed.p) : nullptr;
    A.brick = (want_packed && !use_slab && e->packed_brick) ? 1 : 0;
    A.raw_slab = use_raw_slab ? static_cast<const uint2 *>(e->raw_slab.p) : nullptr;
    A.nx = (uint32_t)e->dims[0]; A.ny = (uint32_t)e->dims[1]; A.nz = (uint32_t)e->dims[2];
    A.V = (int64_t)(e->dims[0] * e->dims[1] * e->dims[2]);
    for (int i = 0; i < 3; i++) A.fov[i] = e->fov[i];
    A.c = e->P.c; A.s = e->P.s;
    A.lin_pc = e->P.linear_phase_cycling; A.quad_pc = e->P.quadratic_phase_cycling;
    A.rf_ph0 = e->rf_ph0;
    A.field_k = e->P.B0 * e->P.timestep_us * 1e-6 * kGamma * kRad2Deg; // monte_carlo.cu:241
    A.timestep_us = e->P.timestep_us;
    A.n_tp = e->P.n_timepoints;
    A.n_scans = (uint32_t)e->P.n_dummy_scan + 1u;
    A.n_spins_global = e->P.n_spins;
    A.n_te = e->n_te;
    A.seed = e->P.seed;
    A.max_iter = e->P.max_iterations;
    A.cross_fov = e->P.cross_fov;
    A.record = e->P.record_trajectory;
    A.one_bits = 0x3f800000u;
    for (uint32_t r = 0; r < 10; r++) A.perm_key[r] = perm_stream_key(e->P.seed) + r * 0x9E3779B9u;
    A.blob = static_cast<const uint8_t *>(e->blob.p);
    A.L = e->L;
    A.scales = static_cast<const float *>(e->scales.p);
    A.n_scales = n_scales;
    A.scale_type = scale_type;
    A.xyz0 = static_cast<const float *>(e->xyz0.p);
    A.m0 = e->has_m0 ? static_cast<const float *>(e->m0.p) : nullptr;
    A.order = e->order_valid ? static_cast<const uint32_t *>(e->order.p) : nullptr;
    A.spin_first = e->spin_first;
    A.n_local = e->n_local;
    A.stage = static_cast<uint4 *>(e->stage.p);
    A.stage_row = (uint32_t)stage_row;
    A.stage_chunks = (uint32_t)stage_chunks;
    A.stage_by_slot = 1; // (re-binned runs, whose order changes between legs, index their rows by the spin: set below)
    A.XYZ1 = record ? static_cast<float *>(e->XYZ1.p) : nullptr;
    A.sums_fx = (E * ns) ? static_cast<unsigned long long *>(e->sums_fx.p) : nullptr;
    A.counters = static_cast<unsigned long long *>(e->counters.p);
    A.trj = e->trj;
    {
        const float *pxy = reinterpret_cast<const float *>(e->blob_h.data() + e->L.pXY);
        for (size_t i = 0; i < ns * ns; i++) A.perm_draws |= (pxy[i] > 0.f && pxy[i] < 1.f) ? 1 : 0;
    }

    const bool stats_on = (flags & SWK_RUN_STATS) != 0;
    const size_t smem_cap = std::min<size_t>(e->smem_optin, 96 * 1024);
    if (mode == SWK_MODE_FAST && (uint64_t)A.V >= (1ull << 32)) return fail(e, SWK_ERR_INVALID, "SWK_MODE_FAST indexes voxels with 32 bits: phantom too large");
    if (mode == SWK_MODE_FAST && std::max(A.nx, std::max(A.ny, A.nz)) > (1u << 20)) return fail(e, SWK_ERR_INVALID, "SWK_MODE_FAST: more than 2^20 voxels along one axis");

    // ---- kernel variants and their shared memory ----
    // COMPAT: [tables] [block sums].  FAST: [tables] [block sums x scales of the block] [scale constants] [event state] [normals, SHARED variant].
    // A run is cut into `parts`: launches over contiguous scale ranges, each with its own kernel variant.
    struct Part {
        uint32_t k_lo, k_hi, group, n_groups;
        unsigned block;
        bool shared;
        size_t fixed, smem;
        walk_fn kern;
    };
    std::vector<Part> parts;
    walk_fn kern_private = nullptr; // FAST: the variant of legs resumed after a re-binning pause (walkers resume at their own rounds)
    bool onewalk = false;           // FAST: one walker per spin for all (gradient / phase-cycling) scales
    size_t multi_chunk = 0;         // ... walked in chunks of this many thread slots (== S unless the run has no per-spin outputs)
    size_t smem_private = 0;
    const size_t bsum_bytes = A.sums_fx ? E * ns * 4 * sizeof(long long) : 0;
    if (mode == SWK_MODE_COMPAT) {
        if (bsum_bytes > smem_cap) return fail(e, SWK_ERR_INVALID, "too many echoes x substrates for the in-kernel ensemble sums");
        A.blob_in_smem = (e->L.bytes + bsum_bytes <= smem_cap) ? 1 : 0;
        Part p{};
        p.k_lo = 0; p.k_hi = n_scales; p.block = kBlock;
        p.smem = (A.blob_in_smem ? e->L.bytes : 0) + bsum_bytes;
        p.kern = stats_on ? walk_compat_kernel<true> : walk_compat_kernel<false>;
        CK(cudaFuncSetAttribute(p.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
        parts.push_back(p);
    } else {
        std::vector<uint8_t> tab;
        uint32_t stride = 0;
        scale_constants(e, scales, n_scales, scale_type, tab, stride);
        if ((rc = ensure(e, e->scale_tab, tab.size())) != SWK_OK) return rc;
        CK(cudaMemcpyAsync(e->scale_tab.p, tab.data(), tab.size(), cudaMemcpyHostToDevice, e->stream));
        CK(cudaStreamSynchronize(e->stream)); // `tab` is a local
        A.scale_tab = static_cast<const uint8_t *>(e->scale_tab.p);
        A.scale_stride = stride;
        const int vox = A.packed ? (use_slab ? VOX_SLAB : VOX_PACKED) : (A.fieldmap ? VOX_SPLIT : VOX_MASK);
        bool gruns = false; // does the timeline hold a run of gradient samples (swk_set_sequence: tl_run >= 2)?
        {
            const uint32_t *run = reinterpret_cast<const uint32_t *>(e->blob_h.data() + e->L.tl_run);
            for (uint32_t i = 0; i < e->L.n_tl; i++) gruns |= run[i] >= 2u;
        }
        { // nominal rounds of one TR for a walker that is never rejected: its segments (walk_fast.cuh advance_walker), each ending at a sync round
            const int32_t *tl_time = reinterpret_cast<const int32_t *>(e->blob_h.data() + e->L.tl_time);
            const uint32_t *tl_run = reinterpret_cast<const uint32_t *>(e->blob_h.data() + e->L.tl_run);
            const uint32_t n_tp = A.n_tp;
            uint32_t r = 0, t = 0;
            auto seg = [&](uint32_t stop) {
                if (stop > t) { r = (r + (stop - t) + kSync - 1) / kSync * kSync; t = stop; }
            };
            for (uint32_t ev = 0; ev < e->L.n_tl;) {
                const uint32_t ev_time = (uint32_t)tl_time[ev];
                if (ev_time >= n_tp) break;
                if (gruns && tl_run[ev] >= 2u) {
                    const uint32_t len = std::min(tl_run[ev], n_tp - ev_time);
                    seg(ev_time);
                    seg(ev_time + len);
                    ev += len;
                } else {
                    seg(ev_time + 1u);
                    ev++;
                }
            }
            seg(n_tp);
            A.tr_period = (r + 2u + n_tp / 64u + kSync - 1) / kSync * kSync;
        }
        const size_t fixed_private = ((bsum_bytes + 15) & ~size_t(15)) + stride + (size_t)ES_FIELDS * 4 * kBlock;
        if (fixed_private > smem_cap) return fail(e, SWK_ERR_INVALID, "too many echoes x substrates for the in-kernel ensemble sums");
        // ONE WALK FOR ALL SCALES (walk_fast.cuh MULTI): gradient and phase-cycling scales do not change the walk, and the reference replays the same
        // random stream for every scale of a spin (kernels.cu:77-88), so one walker per spin carries the magnetisation of every scale (A.mstate).
        size_t fixed_multi = ((bsum_bytes * K + 15) & ~size_t(15)) + stride + (size_t)ES_FIELDS * 4 * kBlock;
        if (gruns && fixed_multi + (size_t)e->L.n_grad * sizeof(float4) + e->L.bytes <= smem_cap) { // pre-multiplied gradient samples in shared memory
            fixed_multi += (size_t)e->L.n_grad * sizeof(float4);
            A.g4_smem = 1;
        }
        onewalk = scale_type != SWK_SCALE_FOV && K > 1 && !record && !(flags & SWK_RUN_NO_ONEWALK) && getenv("SWK_NO_ONEWALK") == nullptr &&
                  fixed_multi <= smem_cap;
        if (onewalk) {
            // magnetisations of every scale between two events: 16 bytes per (scale, walker).  Runs with per-spin outputs hold 16 (E + 1) bytes of staging
            // rows per (scale, walker) anyway; runs without them (ensemble sums only: 1e8 and more spins per GPU) walk chunks of slots one after the other
            // so that this buffer stays bounded
            multi_chunk = (!want_stage && n_slices == 1) ? (size_t)std::min<uint64_t>(S, std::max<uint64_t>(kBlock, ((256ull << 20) / K) / kBlock * kBlock)) : S;
            if (const char *ev = getenv("SWK_MULTI_CHUNK")) multi_chunk = std::min<size_t>(S, (size_t)std::max(1, atoi(ev)) * kBlock); // test hook
            if ((rc = ensure(e, e->mstate, K * multi_chunk * sizeof(uint4))) != SWK_OK) return rc;
            A.n_multi = (uint32_t)K;
            A.mstate = static_cast<uint4 *>(e->mstate.p);
            A.m_first = 0;
            A.m_rows = (uint32_t)multi_chunk;
        }
        // SHARED variant (walk_fast.cuh): 32 spins x G scales per block share the spins' normals: ~45 instead of ~90 instructions per attempt
        // where the launch is issue bound.  Where every attempt fetches a voxel far from the last one (step sigma above ~1 voxel: the small FoV
        // scales) the walk is bound by the gather pipeline instead, the lockstep of a block's scales only costs, and the PRIVATE variant is
        // faster (measured on C2, profiles/README.md).  A run is therefore cut into at most two launches over contiguous scale ranges.
        const size_t nbuf_bytes = 2 * kBatch * 32 * sizeof(float4);
        auto shared_bytes = [&](uint32_t g) { return ((bsum_bytes * g + 15) & ~size_t(15)) + (size_t)stride * g + (size_t)ES_FIELDS * 4 * 32 * g + nbuf_bytes; };
        double sig_thr = 1.25;
        if (const char *ev = getenv("SWK_SHARE_SIGMA")) sig_thr = atof(ev); // tuning knob
        bool share_off = (flags & SWK_RUN_NO_SHARE) || getenv("SWK_NO_SHARE") != nullptr;
        // A voxel table that lives in HBM (no z slab, larger than L2): the launch as a whole runs at the random-access rate of the memory, and what
        // matters is that gather-bound and issue-bound blocks share every SM — one PRIVATE launch over all scales does that (measured on C2's full
        // table: 557 ms against 632 ms for the two-launch cut, profiles/README.md).
        {
            const size_t table_bytes = A.packed ? (use_slab ? e->slab.bytes : e->packed.bytes) : (size_t)A.V * (A.fieldmap ? 5 : 1);
            if (table_bytes > (size_t)100e6 && getenv("SWK_SHARE_SIGMA") == nullptr) share_off = true;
        }
        std::vector<char> want_shared(n_scales, 0);
        {
            const double *sig = reinterpret_cast<const double *>(e->blob_h.data() + e->L.sigma);
            for (uint32_t k = 0; k < n_scales; k++) {
                double sv = 0.;
                for (uint32_t sub = 0; sub < ns; sub++)
                    for (int i = 0; i < 3; i++) sv = std::max(sv, sig[sub] * (double)e->dims[i] / ((double)e->fov[i] * (scale_type == SWK_SCALE_FOV ? (double)scales[k] : 1.)));
                want_shared[k] = !share_off && sv <= sig_thr;
            }
        }
        // [0, cut) one way and [cut, K) the other when the scales are ordered that way (the default list ascends); a mixed order is decided by the majority
        uint32_t cut = 0;
        while (cut < n_scales && want_shared[cut] == want_shared[0]) cut++;
        bool two = cut < n_scales;
        for (uint32_t k = cut; two && k < n_scales; k++) two = want_shared[k] == want_shared[cut];
        if (cut < n_scales && !two) {
            uint32_t n_sh = 0;
            for (uint32_t k = 0; k < n_scales; k++) n_sh += want_shared[k];
            std::fill(want_shared.begin(), want_shared.end(), (char)(2 * n_sh >= n_scales));
            cut = n_scales;
        }
        A.blob_in_smem = 1;
        auto add_part = [&](uint32_t k_lo, uint32_t k_hi, bool sh) {
            Part p{};
            p.k_lo = k_lo; p.k_hi = k_hi;
            uint32_t G = 1;
            if (sh) {
                G = shared_group(k_hi - k_lo);
                if (const char *ev = getenv("SWK_GROUP")) G = (uint32_t)std::max(1, std::min(atoi(ev), (int)std::min<uint32_t>(k_hi - k_lo, SWK_FAST_SHARED_MAXT / 32))); // tuning knob
                while (G > 1 && shared_bytes(G) > smem_cap) G--; // (many echoes x substrates: fewer scales per block)
                if (G < 4) G = 1;                                 // too few walkers per generation of normals to pay for the block barrier
            }
            p.shared = G >= 2;
            p.group = p.shared ? G : 0;
            p.n_groups = p.shared ? (k_hi - k_lo + G - 1) / G : 0;
            p.block = p.shared ? 32 * G : (unsigned)kBlock;
            p.fixed = p.shared ? shared_bytes(G) : fixed_private;
            p.kern = p.shared ? pick_fast<true>(vox, gruns, record, stats_on) : pick_fast<false>(vox, gruns, record, stats_on);
            if (e->L.bytes + p.fixed > smem_cap) A.blob_in_smem = 0;
            parts.push_back(p);
        };
        if (onewalk) { // one launch: a block = kBlock spins, geometry of a single scale
            Part p{};
            p.k_lo = 0; p.k_hi = 1; p.block = (unsigned)kBlock;
            p.fixed = fixed_multi;
            p.kern = pick_fast_multi(vox, gruns, stats_on);
            if (e->L.bytes + p.fixed > smem_cap) A.blob_in_smem = 0;
            parts.push_back(p);
        } else {
            add_part(0, cut, want_shared[0] != 0);
            if (cut < n_scales) add_part(cut, n_scales, want_shared[cut] != 0);
            if (parts.size() == 2 && !parts[0].shared && !parts[1].shared) { // (e.g. too few scales on the shared side)
                parts[0].k_hi = n_scales;
                parts.pop_back();
            }
        }
        for (Part &p : parts) {
            p.smem = (A.blob_in_smem ? e->L.bytes : 0) + p.fixed;
            CK(cudaFuncSetAttribute(p.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
        }
        kern_private = pick_fast<false>(vox, gruns, record, stats_on);
        smem_private = (A.blob_in_smem ? e->L.bytes : 0) + fixed_private;
        CK(cudaFuncSetAttribute(kern_private, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
    }
    // one walk over thread slots [A.j_first, A.j_end): the launches of `parts` (with `fork`, the second part runs beside the first on a second stream)
    uint32_t walk_launches = 0;
    auto launch_walk = [&](cudaStream_t st, bool fork, bool resumed) -> int {
        const uint64_t n = A.j_end - A.j_first;
        if (resumed) { // FAST legs after a re-binning pause: every walker resumes at its own round
            A.k_lo = 0; A.k_hi = n_scales; A.group = 0; A.n_groups = 0;
            const uint64_t grid = ((n + kBlock - 1) / kBlock) * K;
            if (grid > 0x7fffffffull) return fail(e, SWK_ERR_INVALID, "too many spins x scales for one launch");
            kern_private<<<(unsigned)grid, kBlock, smem_private, st>>>(A);
            CK(cudaGetLastError());
            walk_launches++;
            return SWK_OK;
        }
        for (size_t ip = 0; ip < parts.size(); ip++) {
            const Part &p = parts[ip];
            A.k_lo = p.k_lo; A.k_hi = p.k_hi; A.group = p.group; A.n_groups = p.n_groups;
            const uint64_t grid = p.shared ? ((n + 31) / 32) * p.n_groups : ((n + kBlock - 1) / kBlock) * (p.k_hi - p.k_lo);
            if (grid > 0x7fffffffull) return fail(e, SWK_ERR_INVALID, "too many spins x scales for one launch");
            cudaStream_t s2 = st;
            if (fork && ip == 1) { // the second range beside the first
                if (!e->cstream[0]) CK(cudaStreamCreateWithFlags(&e->cstream[0], cudaStreamNonBlocking));
                if (!e->ev_fork) CK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
                if (!e->ev_join) CK(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
                s2 = e->cstream[0];
                CK(cudaEventRecord(e->ev_fork, st));
                CK(cudaStreamWaitEvent(s2, e->ev_fork, 0));
            }
            p.kern<<<(unsigned)grid, p.block, p.smem, s2>>>(A);
            CK(cudaGetLastError());
            walk_launches++;
            if (s2 != st) {
                CK(cudaEventRecord(e->ev_join, s2));
                CK(cudaStreamWaitEvent(st, e->ev_join, 0));
            }
        }
        return SWK_OK;
    };
    // staging rows -> reference layouts, rows [r0, r1) of every scale (no-op when no per-spin output was requested)
    auto unpack = [&](size_t r0, size_t r1, cudaStream_t st) -> cudaError_t {
        if (!e->stage.p || r1 <= r0) return cudaSuccess;
        const size_t total = (r1 - r0) * stage_row * K;
        const unsigned g = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)e->sm_count * 32);
        unpack_rows_kernel<<<g, 256, 0, st>>>(static_cast<const uint4 *>(e->stage.p), (uint32_t)stage_row, (uint32_t)E, S, stage_chunks, r0, r1, (uint32_t)K,
                                             (A.stage_by_slot && e->order_valid) ? static_cast<const uint32_t *>(e->inv_order.p) : nullptr,
                                             static_cast<float *>(e->M1.p), static_cast<uint8_t *>(e->T.p), record ? nullptr : static_cast<float *>(e->XYZ1.p));
        return cudaGetLastError();
    };

    CK(cudaEventRecord(e->ev0, e->stream));
    A.scan_first = 0;
    A.scan_end = A.n_scans;
    if (n_slices == 1) {
        A.j_first = 0;
        A.j_end = (uint32_t)S;
        // ---- long runs (bSSFP: ~1000 TRs): pause at TR boundaries and re-sort the spins by their CURRENT voxel ----
        // The start order keeps the resident spins' voxels inside L2 only while they have not diffused apart: after n steps the
        // cloud has grown by ~2 sigma_vox sqrt(n) voxels per axis.  A pause every (35 / sigma_vox)^2 steps keeps that growth below
        // ~70 voxels (measured on C4: 2.26e11 steps/s at 60 TRs per leg, 2.13e11 at 125, 2.01e11 at 250, 1.65e11 without); it costs one 36-byte state record per walker and one radix sort.  Small FoV scales (sigma_vox > 2) touch the
        // table at random whatever the order and are not re-binned; a z-slab table is cache resident whatever the order.
        uint32_t scans_per_leg = A.n_scans;
        bool per_scale = false;
        if (mode == SWK_MODE_FAST && A.order && !A.record && !(flags & SWK_RUN_NO_REBIN) && A.n_scans > 1 && !onewalk) { // (a MULTI walk is never paused)
            double sig_vox = 0.;
            const double *sig = reinterpret_cast<const double *>(e->blob_h.data() + e->L.sigma);
            for (uint32_t k = 0; k < n_scales; k++)
                for (uint32_t sub = 0; sub < ns; sub++)
                    for (int i = 0; i < 3; i++)
                        sig_vox = std::max(sig_vox, sig[sub] * (double)e->dims[i] / ((double)e->fov[i] * (scale_type == SWK_SCALE_FOV ? (double)scales[k] : 1.)));
            const double n_rb = sig_vox > 0. ? (35. / sig_vox) * (35. / sig_vox) : 1e30;
            const double total = (double)A.n_scans * A.n_tp;
            // the walks differ between scales only when the FoV is scaled; the scale index is the top byte of the re-binning key, so
            // runs with more than 256 scales share one order (that of scale 0: any order is valid, only locality suffers)
            per_scale = scale_type == SWK_SCALE_FOV && K > 1 && K <= 256;
            if (!use_slab && sig_vox <= 2. && total >= 2. * n_rb && (!per_scale || K * S <= (1ull << 29)))
                scans_per_leg = (uint32_t)std::max(1., std::floor(n_rb / A.n_tp));
            if (const char *ev = getenv("SWK_REBIN_SCANS")) scans_per_leg = (uint32_t)std::max(1, atoi(ev)); // tuning knob
        }
        if (scans_per_leg < A.n_scans) {
            const size_t n_state = K * S, n_seg = per_scale ? K : 1, n_sort = n_seg * S;
            if ((rc = ensure(e, e->state_a, n_state * sizeof(uint4))) != SWK_OK || (rc = ensure(e, e->state_b, n_state * sizeof(uint4))) != SWK_OK ||
                (rc = ensure(e, e->state_vox, n_state * sizeof(uint32_t))) != SWK_OK || (rc = ensure(e, e->sort_keys_in, n_sort * 8)) != SWK_OK ||
                (rc = ensure(e, e->sort_keys_out, n_sort * 8)) != SWK_OK || (rc = ensure(e, e->sort_ids, n_sort * 4)) != SWK_OK)
                return rc;
            A.stage_by_slot = 0;
            A.state_a = static_cast<uint4 *>(e->state_a.p);
            A.state_b = static_cast<uint4 *>(e->state_b.p);
            A.state_vox = static_cast<uint32_t *>(e->state_vox.p);
            DevBuf order2; // the re-binned order lives in its own buffer: e->order keeps the start order for the next run
            struct Free { DevBuf &b; ~Free() { release(b); } } free_order2{order2};
            if ((rc = ensure(e, order2, n_sort * sizeof(uint32_t))) != SWK_OK) return rc;
            for (uint32_t s0 = 0; s0 < A.n_scans; s0 += scans_per_leg) {
                A.scan_first = s0;
                A.scan_end = std::min(A.n_scans, s0 + scans_per_leg);
                if ((rc = launch_walk(e->stream, false, s0 != 0)) != SWK_OK) return rc;
                if (A.scan_end == A.n_scans) break;
                rebin_keys_kernel<<<(unsigned)((n_sort + 255) / 256), 256, 0, e->stream>>>(A.state_vox, A.state_b, (uint32_t)S, (uint32_t)n_seg, A.ny, A.nz,
                                                                                         static_cast<uint64_t *>(e->sort_keys_in.p), static_cast<uint32_t *>(e->sort_ids.p));
                CK(cudaGetLastError());
                size_t tmp_bytes = 0;
                CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, static_cast<uint64_t *>(e->sort_keys_in.p), static_cast<uint64_t *>(e->sort_keys_out.p),
                                                   static_cast<uint32_t *>(e->sort_ids.p), static_cast<uint32_t *>(order2.p), (int)n_sort, 0, 64, e->stream));
                if ((rc = ensure(e, e->sort_tmp, std::max(tmp_bytes, e->sort_tmp.bytes))) != SWK_OK) return rc;
                CK(cub::DeviceRadixSort::SortPairs(e->sort_tmp.p, tmp_bytes, static_cast<uint64_t *>(e->sort_keys_in.p), static_cast<uint64_t *>(e->sort_keys_out.p),
                                                   static_cast<uint32_t *>(e->sort_ids.p), static_cast<uint32_t *>(order2.p), (int)n_sort, 0, 64, e->stream));
                A.order = static_cast<const uint32_t *>(order2.p);
                A.order_per_scale = per_scale ? 1 : 0;
                extra_launches += 2;
            }
            CK(unpack(0, S, e->stream));
            CK(cudaStreamSynchronize(e->stream)); // order2 is freed on return
        } else if (onewalk && multi_chunk < S) { // ensemble sums only: chunk after chunk through the same magnetisation buffer
            for (size_t c0 = 0; c0 < S; c0 += multi_chunk) {
                A.j_first = (uint32_t)c0;
                A.j_end = (uint32_t)std::min(S, c0 + multi_chunk);
                A.m_first = A.j_first;
                if ((rc = launch_walk(e->stream, false, false)) != SWK_OK) return rc;
            }
        } else {
            if ((rc = launch_walk(e->stream, true, false)) != SWK_OK) return rc;
            CK(unpack(0, S, e->stream));
        }
    } else {
        // slice i runs on compute stream i & 1 (tails overlap the next slice); its rows are unpacked there and downloaded as soon as they are done
        const bool one_stream = getenv("SWK_ONE_CSTREAM") != nullptr; // tuning knob
        for (int c = 0; c < 2; c++) CK(cudaStreamWaitEvent(e->cstream[c], e->ev0, 0));
        for (uint32_t i = 0; i < n_slices; i++) {
            A.j_first = i * slice_len;
            A.j_end = (uint32_t)std::min<size_t>(S, (size_t)(i + 1) * slice_len);
            cudaStream_t cs = e->cstream[one_stream ? 0 : (i & 1)];
            if ((rc = launch_walk(cs, false, false)) != SWK_OK) return rc;
            CK(unpack(A.j_first, A.j_end, cs)); // the sorted order is slice-major: slots [j_first, j_end) hold exactly the spins [j_first, j_end)
            CK(cudaEventRecord(e->ev_slice[i], cs));
        }
        CK(cudaStreamWaitEvent(e->stream, e->ev_slice[n_slices - 1], 0));
        CK(cudaStreamWaitEvent(e->stream, e->ev_slice[n_slices - 2], 0));
    }
    if (E * ns) {
        const size_t n = K * E * ns * 4;
        sums_to_double_kernel<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(static_cast<const unsigned long long *>(e->sums_fx.p), n, sums);
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(e->ev1, e->stream));
    if (host && n_slices > 1) { // rows [j_first, j_end) of every scale: one strided copy per array and slice
        for (uint32_t i = 0; i < n_slices; i++) {
            const size_t r0 = (size_t)i * slice_len, r1 = std::min<size_t>(S, r0 + slice_len);
            const size_t HS = e->host_rows ? (size_t)e->host_rows : S, h0 = (e->host_rows ? (size_t)e->host_row_first : 0) + r0; // host pitch / first row
            CK(cudaStreamWaitEvent(e->dstream, e->ev_slice[i], 0));
            const size_t rows[3] = {E * 3 * sizeof(float), (size_t)e->trj * 3 * sizeof(float), E};
            char *dst[3] = {reinterpret_cast<char *>(host->M1), reinterpret_cast<char *>(host->XYZ1), reinterpret_cast<char *>(host->T)};
            const char *src[3] = {static_cast<const char *>(e->M1.p), static_cast<const char *>(e->XYZ1.p), static_cast<const char *>(e->T.p)};
            for (int a = 0; a < 3; a++)
                if (dst[a] && src[a] && rows[a])
                    CK(copy_rows(dst[a] + h0 * rows[a], HS * rows[a], src[a] + r0 * rows[a], S * rows[a], (r1 - r0) * rows[a], K, e->mem_pitch, e->dstream));
        }
    }
    tr.mark("launches enqueued");
    unsigned long long cnt[8] = {0}; // pageable destination: this copy blocks until the kernels are done, so it comes after all enqueues
    CK(cudaMemcpyAsync(cnt, e->counters.p, sizeof cnt, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream)); // ≙ the device sync after the launch (monte_carlo.cu:333)
    tr.mark("kernels done");
    if (host && n_slices > 1) CK(cudaStreamSynchronize(e->dstream));
    tr.mark("downloads done");
    float ms = 0.f, ms_all = 0.f;
    CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    CK(cudaEventElapsedTime(&ms_all, e->evA, e->ev1));

    e->n_scales = n_scales;
    e->last_used_slab = use_slab;
    e->out_flags = flags;
    e->last_sums = (E * ns) ? sums : nullptr;
    e->last_slices = n_slices;
    swk_stats &st = e->stats;
    st = swk_stats{};
    const uint64_t nominal = (uint64_t)S * K * A.n_scans * A.n_tp;
    st.steps = stats_on ? cnt[0] : nominal;
    st.mask_gathers = cnt[1];
    st.field_gathers = cnt[2];
    st.rejects = cnt[3];
    st.lost = cnt[4];
    st.kernel_ms = ms;
    st.device_ms = ms_all;
    st.n_launches = walk_launches + extra_launches + (e->stage.p ? n_slices : 0) + ((E * ns) ? 1 : 0);
    return SWK_OK;
}

extern "C" {

int swk_run_device(swk_engine *e, const float *scales, uint32_t n_scales, int scale_type, int mode, int flags, double *d_sums)
{
    return run_impl(e, scales, n_scales, scale_type, mode, flags, d_sums, 1, nullptr);
}

int swk_download(swk_engine *e, float *M1, float *XYZ1, uint8_t *T)
{
    if (!e) return SWK_ERR_INVALID;
    if (e->n_scales == 0) return fail(e, SWK_ERR_STATE, "swk_download: nothing has been run");
    if ((M1 && !e->M1.p) || (XYZ1 && !e->XYZ1.p) || (T && !e->T.p)) return fail(e, SWK_ERR_STATE, "swk_download: output was not requested in the run flags");
    CK(cudaSetDevice(e->device));
    if (!e->host_rows) {
        if (M1) CK(cudaMemcpyAsync(M1, e->M1.p, e->M1.bytes, cudaMemcpyDeviceToHost, e->stream));
        if (XYZ1) CK(cudaMemcpyAsync(XYZ1, e->XYZ1.p, e->XYZ1.bytes, cudaMemcpyDeviceToHost, e->stream));
        if (T) CK(cudaMemcpyAsync(T, e->T.p, e->T.bytes, cudaMemcpyDeviceToHost, e->stream));
    } else { // rows [host_row_first, +n_local) of every scale of the caller's global arrays
        const size_t K = e->n_scales, S = e->n_local, HS = (size_t)e->host_rows, h0 = (size_t)e->host_row_first;
        const size_t rows[3] = {(size_t)e->n_te * 3 * sizeof(float), (size_t)e->trj * 3 * sizeof(float), (size_t)e->n_te};
        char *dst[3] = {reinterpret_cast<char *>(M1), reinterpret_cast<char *>(XYZ1), reinterpret_cast<char *>(T)};
        const char *src[3] = {static_cast<const char *>(e->M1.p), static_cast<const char *>(e->XYZ1.p), static_cast<const char *>(e->T.p)};
        for (int a = 0; a < 3; a++)
            if (dst[a] && rows[a]) CK(copy_rows(dst[a] + h0 * rows[a], HS * rows[a], src[a], S * rows[a], S * rows[a], K, e->mem_pitch, e->stream));
    }
    CK(cudaStreamSynchronize(e->stream));
    return SWK_OK;
}

int swk_get_sums(swk_engine *e, double *sums)
{
    if (!e || !sums) return SWK_ERR_INVALID;
    if (e->n_scales == 0) return fail(e, SWK_ERR_STATE, "swk_get_sums: nothing has been run");
    const size_t n = (size_t)e->n_scales * e->n_te * e->P.n_substrate * 4;
    if (n == 0) return SWK_OK;
    if (!e->last_sums) return fail(e, SWK_ERR_STATE, "swk_get_sums: no sums were accumulated");
    CK(cudaSetDevice(e->device));
    CK(cudaMemcpyAsync(sums, e->last_sums, n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SWK_OK;
}

int swk_get_stats(swk_engine *e, swk_stats *out)
{
    if (!e || !out) return SWK_ERR_INVALID;
    *out = e->stats;
    return SWK_OK;
}

int swk_run(swk_engine *e, const float *XYZ0, const float *M0, uint32_t spin_first, uint32_t n_local, const float *scales, uint32_t n_scales,
            int scale_type, int mode, float *M1, float *XYZ1, uint8_t *T, double *sums, swk_stats *stats)
{
    if (!e) return SWK_ERR_INVALID;
    int rc;
    Trace tr;
    if ((rc = swk_set_spins(e, XYZ0, M0, spin_first, n_local)) != SWK_OK) return rc;
    tr.mark("set_spins (H2D)");
    const int flags = (M1 ? SWK_OUT_M1 : 0) | (XYZ1 ? SWK_OUT_XYZ1 : 0) | (T ? SWK_OUT_T : 0) | (stats ? SWK_RUN_STATS : 0);
    // Large runs are cut into <= 16 slices of >= 2^19 spins so that the device-to-host copy of slice i overlaps the walk of slice
    // i+1; the download of the LAST slice is what stays exposed (measured on C2, 1e7 spins, 12.5 GB out: 1462 / 1412 / 1387 ms per pass
    // end to end with 4 / 8 / 16 slices against 1329 ms of kernels, profiles/README.md).
    uint32_t n_slices = (M1 || XYZ1 || T) ? std::min<uint32_t>(16u, std::max<uint32_t>(1u, n_local >> 19)) : 1u;
    // Long runs (bSSFP: 220 200 steps per walker) stay in one piece: their download is < 1 % of the walk (a walker's 25 output bytes
    // cost what ~160 steps cost), and only an unsliced run pauses for re-binning (run_impl) — C4: 1.56e11 sliced vs 2.15e11 spin-steps/s.
    if (e->has_sequence && !e->P.record_trajectory && (uint64_t)e->P.n_timepoints * (uint64_t)(e->P.n_dummy_scan + 1) >= 16000u) n_slices = 1;
    if (const char *ev = getenv("SWK_SLICES")) n_slices = std::max(1, atoi(ev)); // tuning knob
    HostOut host;
    host.M1 = M1; host.XYZ1 = XYZ1; host.T = T;
    if ((rc = run_impl(e, scales, n_scales, scale_type, mode, flags, nullptr, n_slices, &host)) != SWK_OK) return rc;
    if (e->last_slices == 1 && (rc = swk_download(e, M1, XYZ1, T)) != SWK_OK) return rc; // small runs: plain download
    if (sums && (rc = swk_get_sums(e, sums)) != SWK_OK) return rc;
    if (stats) *stats = e->stats;
    return SWK_OK;
}

int swk_set_host_rows(swk_engine *e, uint64_t n_rows_total, uint64_t row_first)
{
    if (!e) return SWK_ERR_INVALID;
    e->host_rows = n_rows_total;
    e->host_row_first = n_rows_total ? row_first : 0;
    return SWK_OK;
}

int swk_alloc_pinned(void **ptr, size_t bytes)
{
    if (!ptr) return SWK_ERR_INVALID;
    *ptr = nullptr;
    return cudaHostAlloc(ptr, bytes, cudaHostAllocPortable) == cudaSuccess ? SWK_OK : SWK_ERR_MEMORY;
}

void swk_free_pinned(void *ptr)
{
    if (ptr) cudaFreeHost(ptr);
}

int swk_probe_gather(swk_engine *e, uint32_t threads_per_sm, uint32_t iters, double *gathers_per_s, uint64_t *table_bytes)
{
    if (!e || !gathers_per_s) return SWK_ERR_INVALID;
    if (!e->has_phantom) return fail(e, SWK_ERR_STATE, "swk_probe_gather: no phantom (swk_set_phantom)");
    if (iters == 0 || threads_per_sm == 0) return fail(e, SWK_ERR_INVALID, "swk_probe_gather: iters and threads_per_sm must be positive");
    CK(cudaSetDevice(e->device));
    const DevBuf &t = (e->last_used_slab && e->slab_valid) ? e->slab : ((e->packed_valid && e->packed.p) ? e->packed : (e->fieldmap.p ? e->fieldmap : e->mask));
    const uint32_t n_words = (uint32_t)std::min<size_t>(t.bytes / 4, 0xffffffffu);
    if (n_words == 0) return fail(e, SWK_ERR_STATE, "swk_probe_gather: voxel table is empty");
    int rc;
    if ((rc = ensure(e, e->counters, 8 * sizeof(unsigned long long))) != SWK_OK) return rc;
    const unsigned grid = (unsigned)e->sm_count * std::max(1u, (threads_per_sm + 255u) / 256u);
    uint32_t *sink = static_cast<uint32_t *>(e->counters.p);
    gather_probe_kernel<<<grid, 256, 0, e->stream>>>(static_cast<const uint32_t *>(t.p), n_words, std::max(1u, iters / 8), sink); // warm-up
    CK(cudaEventRecord(e->ev0, e->stream));
    gather_probe_kernel<<<grid, 256, 0, e->stream>>>(static_cast<const uint32_t *>(t.p), n_words, iters, sink);
    CK(cudaEventRecord(e->ev1, e->stream));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    *gathers_per_s = (double)grid * 256.0 * iters / (ms * 1e-3);
    if (table_bytes) *table_bytes = (uint64_t)n_words * 4;
    return SWK_OK;
}

int swk_debug_rng(swk_engine *e, int which, const uint32_t *in, uint32_t n, uint32_t *out)
{
    if (!e) return SWK_ERR_INVALID;
    if (!in || !out || n == 0 || which < 0 || which > 2) return fail(e, SWK_ERR_INVALID, "swk_debug_rng: bad arguments");
    CK(cudaSetDevice(e->device));
    DevBuf d_in, d_out;
    struct Free { DevBuf &a, &b; ~Free() { release(a); release(b); } } guard{d_in, d_out};
    int rc;
    if ((rc = ensure(e, d_in, (size_t)n * 16)) != SWK_OK || (rc = ensure(e, d_out, (size_t)n * 32)) != SWK_OK) return rc;
    CK(cudaMemcpyAsync(d_in.p, in, (size_t)n * 16, cudaMemcpyHostToDevice, e->stream));
    debug_rng_kernel<<<(n + 255) / 256, 256, 0, e->stream>>>(which, static_cast<const uint32_t *>(d_in.p), n, 0x3f800000u, static_cast<uint32_t *>(d_out.p));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_out.p, (size_t)n * 32, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SWK_OK;
}

void *swk_stream(swk_engine *e) { return e ? (void *)e->stream : nullptr; }
double *swk_device_sums(swk_engine *e) { return e ? e->last_sums : nullptr; }
uint64_t swk_device_bytes(const swk_engine *e)
{
    if (!e) return 0;
    uint64_t n = 0;
    for (const DevBuf *b : {&e->mask, &e->fieldmap, &e->packed, &e->blob, &e->xyz0, &e->m0, &e->order, &e->scales, &e->M1, &e->XYZ1, &e->T, &e->sums, &e->counters,
                            &e->sort_keys_in, &e->sort_keys_out, &e->sort_ids, &e->sort_tmp, &e->state_a, &e->state_b, &e->state_vox, &e->mstate, &e->raw_slab, &e->slab, &e->stage, &e->sums_fx, &e->scale_tab, &e->inv_order})
        n += b->bytes;
    return n;
}

} // extern "C"

// ------------------------------------------------------------------------------------------------------------------------
// Phantom generator (include/spinwalk_phantom.h)
// ------------------------------------------------------------------------------------------------------------------------
#include "phantom.cuh"
#include "phantom_mesh.cuh"

namespace {
thread_local std::string g_phantom_error;
int phantom_fail(int code, const std::string &msg)
{
    g_phantom_error = msg;
    return code;
}
void fill_stats(swk_phantom_stats *st, const swk_phantom_spec &sp, const std::vector<swk::phantom::Shape> &shapes, float place_ms, const swk::phantom::FillResult &fr)
{
    if (!st) return;
    const double V = double(sp.resolution) * sp.resolution * sp.resolution;
    st->n_shapes = uint32_t(shapes.size());
    st->volume_fraction = float(double(fr.ones) * 100.0 / V); // ≙ accumulate(mask) * 100.0 / size (phantom_cylinder.cpp:270)
    st->place_ms = place_ms;
    st->kernel_ms = fr.kernel_ms;
    st->n_launches = fr.n_launches;
    st->exact_columns = fr.exact_columns;
}
} // namespace

extern "C" {

const char *swk_phantom_last_error(void) { return g_phantom_error.c_str(); }

int swk_phantom_shapes(const swk_phantom_spec *spec, float *shapes, uint32_t cap, uint32_t *n_shapes)
{
    if (!spec || !n_shapes) return phantom_fail(SWK_ERR_INVALID, "swk_phantom_shapes: spec and n_shapes are mandatory");
    std::vector<swk::phantom::Shape> placed;
    float ms = 0.f;
    std::string err;
    const int rc = swk::phantom::place(*spec, placed, ms, err);
    if (rc != SWK_OK) return phantom_fail(rc, err);
    *n_shapes = uint32_t(placed.size());
    for (size_t i = 0; shapes && i < placed.size() && i < cap; i++) {
        shapes[4 * i + 0] = placed[i].x;
        shapes[4 * i + 1] = placed[i].y;
        shapes[4 * i + 2] = placed[i].z;
        shapes[4 * i + 3] = placed[i].r;
    }
    return SWK_OK;
}

int swk_phantom_generate(int device, const swk_phantom_spec *spec, uint8_t *mask, float *fieldmap_T, int on_device, swk_phantom_stats *stats)
{
    if (!spec || !mask) return phantom_fail(SWK_ERR_INVALID, "swk_phantom_generate: spec and mask are mandatory");
    const bool calc = swk::phantom::wants_fieldmap(*spec);
    if (calc && !fieldmap_T) return phantom_fail(SWK_ERR_INVALID, "swk_phantom_generate: oxy_level >= 0 needs a fieldmap buffer");
    std::vector<swk::phantom::Shape> placed;
    float place_ms = 0.f;
    std::string err;
    int rc = swk::phantom::place(*spec, placed, place_ms, err);
    if (rc != SWK_OK) return phantom_fail(rc, err);
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev)
        return phantom_fail(SWK_ERR_CUDA, "swk_phantom_generate: no usable CUDA device (this library has no CPU path for the voxel fill)");
    if (cudaSetDevice(device) != cudaSuccess) return phantom_fail(SWK_ERR_CUDA, "cudaSetDevice failed");
    int sm = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
    const size_t V = size_t(spec->resolution) * spec->resolution * spec->resolution;
    uint8_t *d_mask = mask;
    float *d_field = calc ? fieldmap_T : nullptr;
    if (!on_device) {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if (V * (calc ? 5 : 1) > free_b) return phantom_fail(SWK_ERR_MEMORY, "swk_phantom_generate: phantom does not fit in free device memory");
        d_mask = nullptr;
        d_field = nullptr;
        if (cudaMalloc(&d_mask, V) != cudaSuccess || (calc && cudaMalloc(&d_field, V * sizeof(float)) != cudaSuccess)) {
            if (d_mask) cudaFree(d_mask);
            return phantom_fail(SWK_ERR_MEMORY, "swk_phantom_generate: cudaMalloc failed");
        }
    } else if ((reinterpret_cast<uintptr_t>(mask) & 15) || (calc && (reinterpret_cast<uintptr_t>(fieldmap_T) & 15)))
        return phantom_fail(SWK_ERR_INVALID, "swk_phantom_generate: device buffers must be 16-byte aligned");
    swk::phantom::FillResult fr;
    rc = swk::phantom::fill_device(*spec, placed, d_mask, d_field, nullptr, sm, fr);
    if (rc == SWK_OK && !on_device) {
        // (a double-buffered copy through page-locked staging was measured slower than the driver's own pageable path: 1.54 s vs 1.31 s for 5 GB)
        if (cudaMemcpy(mask, d_mask, V, cudaMemcpyDeviceToHost) != cudaSuccess ||
            (calc && cudaMemcpy(fieldmap_T, d_field, V * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess)) {
            rc = SWK_ERR_CUDA;
            fr.error = "device-to-host copy of the phantom failed";
        }
    }
    if (!on_device) {
        cudaFree(d_mask);
        if (d_field) cudaFree(d_field);
    }
    if (rc != SWK_OK) return phantom_fail(rc, fr.error);
    fill_stats(stats, *spec, placed, place_ms, fr);
    return SWK_OK;
}

int swk_phantom_mesh(int device, float fov_um, uint64_t resolution, const double *vertices, uint64_t n_vertices, const uint64_t *faces, uint64_t n_faces,
                     uint8_t *mask, int on_device, swk_phantom_stats *stats)
{
    if (!mask || (n_vertices && !vertices) || (n_faces && !faces)) return phantom_fail(SWK_ERR_INVALID, "swk_phantom_mesh: vertices, faces and mask are mandatory");
    if (!(fov_um > 0.f) || resolution == 0) return phantom_fail(SWK_ERR_INVALID, "FOV or resolution is not set"); // phantom_base.cpp:110-114
    if (resolution > 2048) return phantom_fail(SWK_ERR_INVALID, "resolution above 2048 is not supported");
    if (n_faces > 0xffffffffull) return phantom_fail(SWK_ERR_INVALID, "more than 2^32 triangles");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev)
        return phantom_fail(SWK_ERR_CUDA, "swk_phantom_mesh: no usable CUDA device (this library has no CPU path for the voxel fill)");
    if (cudaSetDevice(device) != cudaSuccess) return phantom_fail(SWK_ERR_CUDA, "cudaSetDevice failed");
    const size_t V = size_t(resolution) * resolution * resolution;
    uint8_t *d_mask = mask;
    if (!on_device && cudaMalloc(&d_mask, V) != cudaSuccess) return phantom_fail(SWK_ERR_MEMORY, "swk_phantom_mesh: cudaMalloc failed");
    swk::phantom::MeshResult mr;
    int rc = swk::phantom::fill_mesh_device(fov_um, uint32_t(resolution), vertices, n_vertices, faces, n_faces, d_mask, nullptr, mr);
    if (rc == SWK_OK && !on_device && cudaMemcpy(mask, d_mask, V, cudaMemcpyDeviceToHost) != cudaSuccess) {
        rc = SWK_ERR_CUDA;
        mr.error = "device-to-host copy of the mask failed";
    }
    if (!on_device) cudaFree(d_mask);
    if (rc != SWK_OK) return phantom_fail(rc, mr.error);
    if (stats) {
        stats->n_shapes = uint32_t(n_faces);
        stats->volume_fraction = float(double(mr.ones) * 100.0 / double(V));
        stats->place_ms = mr.prep_ms; // host: mesh transform, leaf boxes, row binning
        stats->kernel_ms = mr.kernel_ms;
        stats->n_launches = 1;
        stats->exact_columns = 0;
    }
    return SWK_OK;
}

int swk_generate_phantom(swk_engine *e, const swk_phantom_spec *spec, swk_phantom_stats *stats)
{
    if (!e) return SWK_ERR_INVALID;
    if (!spec) return fail(e, SWK_ERR_INVALID, "swk_generate_phantom: spec is NULL");
    std::vector<swk::phantom::Shape> placed;
    float place_ms = 0.f;
    std::string err;
    int rc = swk::phantom::place(*spec, placed, place_ms, err);
    if (rc != SWK_OK) return fail(e, rc, "swk_generate_phantom: " + err);
    CK(cudaSetDevice(e->device));
    const bool calc = swk::phantom::wants_fieldmap(*spec);
    const size_t V = size_t(spec->resolution) * spec->resolution * spec->resolution;
    e->has_phantom = false;
    e->order_valid = false;
    release(e->mask);
    release(e->fieldmap);
    release(e->packed);
    e->packed_valid = false;
    release(e->slab);
    e->slab_valid = false; e->raw_slab_valid = false;
    e->z_invariant = -1;
    if ((rc = ensure(e, e->mask, V)) != SWK_OK) return rc;
    if (calc && (rc = ensure(e, e->fieldmap, V * sizeof(float))) != SWK_OK) return rc;
    swk::phantom::FillResult fr;
    rc = swk::phantom::fill_device(*spec, placed, static_cast<uint8_t *>(e->mask.p), calc ? static_cast<float *>(e->fieldmap.p) : nullptr, e->stream, e->sm_count, fr);
    if (rc != SWK_OK) return fail(e, rc, "swk_generate_phantom: " + fr.error);
    e->mask_substrates = fr.ones ? 2 : 1; // max(mask) + 1 (monte_carlo.cu:113)
    const float fov_m = spec->fov_um * 1e-6f; // phantom_base.cpp:63
    for (int i = 0; i < 3; i++) { e->dims[i] = spec->resolution; e->fov[i] = fov_m; }
    e->has_phantom = true;
    fill_stats(stats, *spec, placed, place_ms, fr);
    return SWK_OK;
}

int swk_get_phantom(swk_engine *e, uint8_t *mask, float *fieldmap_T)
{
    if (!e) return SWK_ERR_INVALID;
    if (!e->has_phantom) return fail(e, SWK_ERR_STATE, "swk_get_phantom: no phantom");
    CK(cudaSetDevice(e->device));
    const size_t V = size_t(e->dims[0]) * e->dims[1] * e->dims[2];
    if (mask) CK(cudaMemcpyAsync(mask, e->mask.p, V, cudaMemcpyDeviceToHost, e->stream));
    if (fieldmap_T) {
        if (!e->fieldmap.p) return fail(e, SWK_ERR_STATE, "swk_get_phantom: the phantom has no field map");
        CK(cudaMemcpyAsync(fieldmap_T, e->fieldmap.p, V * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    }
    CK(cudaStreamSynchronize(e->stream));
    return SWK_OK;
}

} // extern "C"
