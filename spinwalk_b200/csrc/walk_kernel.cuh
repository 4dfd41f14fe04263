// spinwalk_b200/csrc/walk_kernel.cuh — the per-spin Monte-Carlo time loop, hand-written for sm_100a.
//
// Replaces (not ports) the reference's  sim::cu_sim / sim::sim  (src/sim/kernels.cu:56-233).
// Differences in STRUCTURE (results are the reference's):
//   * one launch covers ALL scales (blockIdx -> (chunk of spins, scale)); the reference launches and
//     device-syncs once per scale (monte_carlo.cu:273-337);
//   * the four per-step table probes (kernels.cu:175,181,190,202) are replaced by a merged, sorted
//     event timeline staged in shared memory: the inner loop walks to the next event time with no
//     event tests at all, and events are handled warp-converged;
//   * spin state (position, magnetisation, RNG, accumulated phase) lives in registers for the whole
//     loop; the reference stores the position to global memory every step (kernels.cu:220-221);
//   * FoV / gradient / phase-cycling scaling happens in-kernel with the same FP32 products the
//     reference forms on the host (monte_carlo.cu:278-280,288-290,303);
//   * the fieldmap stays in Tesla; the Tesla->degree/step factor (monte_carlo.cu:241-244) is applied
//     at the gather with the same FP32 product;
//   * per-(scale, echo, substrate) ensemble sums are reduced in-kernel (warp shuffle -> shared -> one
//     FP64 global atomic per block), which the reference leaves to post-processing.
//
// This file holds the shared types / helpers and the SWK_MODE_COMPAT kernel: minstd_rand + erfcinvf normal +
// FP64 metres — the reference CUDA build's walk bit for bit (same device, same libdevice erfcinvf).
// SWK_MODE_FAST (the product path) is walk_fast.cuh; it shares nothing with this kernel but the helpers.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/spinwalk_engine.h"

namespace swk {

constexpr double kDeg2Rad = 0.0174532925199433; // simulation_parameters.cuh:24
constexpr double kRad2Deg = 57.2957795130823;   // simulation_parameters.cuh:25
constexpr double kGamma   = 267515315.;         // definitions.h:20

constexpr int kBlock = 256;

enum : uint32_t { EV_DEPH = 1u, EV_GRAD = 2u, EV_RF = 4u, EV_ECHO = 8u };

// Offsets (bytes) into the sequence blob; the blob is copied to shared memory when it fits.
struct BlobLayout {
    uint32_t bytes;
    uint32_t tl_time, tl_mask, n_tl;        // merged event timeline: int32 time, uint32 mask
    uint32_t tl_run;                        // uint32 per entry: length of the run of gradient-only entries at consecutive timepoints starting here (else 0)
    uint32_t rf_s, rf_c, rf_ph, n_rf;       // float sin/cos of flip angle, phase (deg); entry 0 unused
    uint32_t deph_deg, n_deph;              // float
    uint32_t gx, gy, gz, n_grad;            // float mT/m (unscaled)
    uint32_t sigma, T1s, T2s, pXY, n_sub;   // double sigma[n_sub] (m); float T1,T2 (s); float pXY[n_sub^2]
};

struct WalkArgs {
    // phantom
    const uint8_t *mask;
    const float   *fieldmap; // Tesla at 1 T, or nullptr
    const uint32_t *packed;  // FAST mode: packed voxel words (engine.cu pack_word), or nullptr
    int32_t  brick;          // the packed volume is stored in 2 x 2 x 4 bricks (walk_fast.cuh table_index)
    const uint2 *raw_slab;   // COMPAT mode, phantom invariant along z: (substrate id, FP32 field bits) of one z plane [nx][ny], or nullptr
    uint32_t nx, ny, nz;
    int64_t  V;
    float    fov[3];         // metres (held as float like the reference, monte_carlo.cuh:37)
    // sequence scalars
    float    c, s, lin_pc, quad_pc, rf_ph0, field_k;
    int32_t  timestep_us;
    uint32_t n_tp, n_scans, n_spins_global, n_te;
    uint64_t seed, max_iter;
    int32_t  cross_fov, record;
    uint32_t one_bits;       // 0x3f800000 (FAST mode: a constant the compiler must keep in a register, walk_fast.cuh and_or)
    uint32_t perm_key[10];   // FAST mode: round keys of the permeability stream (walk_fast.cuh philox2x32_10), key + r * 0x9E3779B9
    // sequence tables
    const uint8_t *blob;
    BlobLayout L;
    int32_t  blob_in_smem;
    // scales
    const float *scales;
    uint32_t n_scales;
    int32_t  scale_type;
    // spins of this shard
    const float *xyz0;       // [n_local][3] unscaled metres
    const float *m0;         // [n_local][3] or nullptr => (0,0,1)
    const uint32_t *order;   // nullptr, or thread j simulates local spin order[j] (locality sort)
    uint32_t spin_first, n_local;
    // long runs (many TRs) are paused at TR boundaries to re-sort the spins by their CURRENT voxel (FAST mode, engine.cu run_impl)
    uint32_t scan_first, scan_end; // this launch simulates scans [scan_first, scan_end) of n_scans
    int32_t  order_per_scale;      // order holds one permutation per scale: [n_scales][n_local]
    uint4   *state_a, *state_b;    // [n_scales][n_local]: (p0, p1, p2, RNG block counter), (Mx, My, Mz, substrate | lost << 8)
    uint32_t *state_vox;           // [n_scales][n_local]: linear voxel index at the pause (sort key of the next launch)
    uint32_t j_first, j_end; // this launch simulates thread slots [j_first, j_end) of the shard (pipelined host runs launch slices)
    // FAST mode: per-scale constants (walk_fast.cuh ScaleConst + sigma table), and how blocks are cut
    const uint8_t *scale_tab;
    uint32_t scale_stride;   // bytes per scale record (multiple of 16)
    uint32_t k_lo, k_hi;     // this launch walks scales [k_lo, k_hi)
    uint32_t group, n_groups; // SHARED variant: scales per block (block = 32 x group threads), groups of this launch per spin chunk
    int32_t  perm_draws;     // some 0 < P_XY < 1: permeability uniforms are needed
    uint32_t tr_period;      // FAST mode, multi-TR runs: nominal rounds per TR (walk_fast.cuh: re-synchronisation at TR boundaries)
    // FAST mode, ONE WALK FOR ALL SCALES (walk_fast.cuh MULTI): when the scales act on the gradients or on the phase cycling, every scale of a spin
    // walks the same path (the reference re-seeds seed+spin per scale, kernels.cu:77-88); one walker per spin then carries n_multi magnetisations
    uint32_t n_multi;        // 0: off, else the number of scales
    uint4   *mstate;         // [n_multi][m_rows], row = thread slot - m_first: (Mx, My, Mz, -) of every scale between two sequence events
    uint32_t m_first, m_rows; // (runs without per-spin outputs walk chunks of m_rows slots one after the other: bounded memory, engine.cu)
    int32_t  g4_smem;        // MULTI kernels with gradient runs: the block keeps (gx, gy, gz) x degrees-per-unit of every gradient sample in shared memory
    // outputs (any may be nullptr)
    // Per-spin results go to STAGING ROWS, one per (scale, local spin): n_te echo slots (Mx, My, Mz, tissue) and one slot for the final
    // position, 16 bytes each — a thread's scattered result write is whole aligned 16/32-byte pieces instead of three partial-sector
    // stores into three arrays; unpack_rows_kernel (engine.cu) streams the rows into the reference layouts below.
    uint4   *stage;          // [K][stage_chunks][stage_row][32]: per warp-sized chunk of rows, structure of arrays (walk_fast.cuh Geo::stage)
    uint32_t stage_row;      // n_te + 1
    uint32_t stage_chunks;   // ceil(n_local / 32)
    int32_t  stage_by_slot;  // rows are indexed by the thread slot (coalesced writes; un-permuted by unpack_rows_kernel), else by the local spin
    float   *XYZ1;           // [K][n_local][trj][3]: written directly only when trajectories are recorded
    unsigned long long *sums_fx; // [K][E][n_sub][4]: sum Mx, My, Mz in fixed point (kSumScale), count
    unsigned long long *counters; // [5]: steps, mask_gathers, field_gathers, rejects, lost
    uint64_t trj;
};

// ------------------------------------------------------------------------------------------------
// magnetisation helpers (kernels.cuh:82-219); FP32 like the reference, contraction left to nvcc
// exactly as in the reference build.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void xrot(float s, float c, const float *m0, float *m1)
{
    m1[0] = m0[0];
    m1[1] = c * m0[1] - s * m0[2];
    m1[2] = s * m0[1] + c * m0[2];
}
__device__ __forceinline__ void yrot(float s, float c, const float *m0, float *m1)
{
    m1[0] = c * m0[0] + s * m0[2];
    m1[1] = m0[1];
    m1[2] = -s * m0[0] + c * m0[2];
}
__device__ __forceinline__ void zrot(float s, float c, const float *m0, float *m1)
{
    m1[0] = c * m0[0] - s * m0[1];
    m1[1] = s * m0[0] + c * m0[1];
    m1[2] = m0[2];
}
// kernels.cuh:160-195
__device__ __forceinline__ void xrot_withphase(float s, float c, float ph, const float *m0, float *m1)
{
    if (ph == 0.0f) { xrot(s, c, m0, m1); return; }
    if (ph == 180.0) { xrot(-s, c, m0, m1); return; }
    if (ph == 90.0) { yrot(s, c, m0, m1); return; }
    if (ph == -90.0 || ph == 270.0) { yrot(-s, c, m0, m1); return; }
    float t[3];
    float sp = sinf(ph * kDeg2Rad), cp = cosf(ph * kDeg2Rad);
    zrot(-sp, cp, m0, m1);
    xrot(s, c, m1, t);
    zrot(sp, cp, t, m1);
}
// kernels.cu:45-52 + kernels.cuh:147-152,214-219.  m is updated in place (m0 -> m1 -> m0 of the reference).
__device__ __forceinline__ void dephase_relax(float *m, float acc_phase_deg, float T1, float T2, float dt_s)
{
    float r[3];
    float sp = sinf(acc_phase_deg * kDeg2Rad), cp = cosf(acc_phase_deg * kDeg2Rad);
    zrot(sp, cp, m, r);
    if (T1 >= 0 && T2 >= 0) {
        float e1 = expf(-dt_s / T1), e2 = expf(-dt_s / T2);
        r[0] = r[0] * e2;
        r[1] = r[1] * e2;
        r[2] = 1. + e1 * (r[2] - 1.);
    }
    m[0] = r[0]; m[1] = r[1]; m[2] = r[2];
}

// the same with the relaxation factors e1 = exp(-dt / T1), e2 = exp(-dt / T2) computed by the caller (they do not depend on the scale: MULTI kernels)
__device__ __forceinline__ void dephase_relax_pre(float *m, float acc_phase_deg, bool relax, float e1, float e2)
{
    float r[3];
    float sp = sinf(acc_phase_deg * kDeg2Rad), cp = cosf(acc_phase_deg * kDeg2Rad);
    zrot(sp, cp, m, r);
    if (relax) {
        r[0] = r[0] * e2;
        r[1] = r[1] * e2;
        r[2] = 1. + e1 * (r[2] - 1.);
    }
    m[0] = r[0]; m[1] = r[1]; m[2] = r[2];
}

// ------------------------------------------------------------------------------------------------
// RNG, compat flavour: thrust::minstd_rand + normal_distribution_nvcc + uniform_real_distribution
// (thrust/random/detail/{linear_congruential_engine.inl,linear_congruential_engine_discard.h,
//  normal_distribution_base.h,uniform_real_distribution.inl}).  Same integers, same FP32 operations
// as the reference's SASS (FFMA u*S1+S2, FADD p+p, inline erfcinvf, FMUL by -/+sqrt2).
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kMinstdA = 48271u, kMinstdM = 2147483647u;

__device__ __forceinline__ uint32_t minstd_mulmod(uint32_t a, uint32_t b)
{ // a*b mod (2^31-1) by Mersenne folding (a,b < 2^31 so the product is < 2^62)
    uint64_t p = (uint64_t)a * b;
    uint32_t r = (uint32_t)(p & kMinstdM) + (uint32_t)(p >> 31);
    r = (r & kMinstdM) + (r >> 31);
    return r == kMinstdM ? 0u : r;
}
__device__ __forceinline__ uint32_t minstd_next(uint32_t &x)
{
    x = minstd_mulmod(x, kMinstdA);
    return x;
}
__device__ inline uint32_t minstd_init(uint64_t seed_plus_spin)
{
    uint32_t x = (uint32_t)seed_plus_spin % kMinstdM; // ctor truncates to uint32, then mod m, 0 -> 1
    if (x == 0) x = 1;
    uint32_t mult = kMinstdA, acc = 1;                // discard(z): x <- a^z x mod m
    for (uint64_t z = seed_plus_spin; z > 0; z >>= 1) {
        if (z & 1) acc = minstd_mulmod(acc, mult);
        mult = minstd_mulmod(mult, mult);
    }
    return minstd_mulmod(acc, x);
}
__device__ __forceinline__ float minstd_normal(uint32_t &x)
{
    const uint32_t range = 2147483645u;
    const float S1 = 4.656612873077392578125e-10f;  // float(1/range) == 2^-31
    const float S2 = 2.3283064365386962890625e-10f; // S1/2
    float S3 = -1.41421353816986083984375f;         // float(-sqrt 2)
    uint32_t u = minstd_next(x) - 1u;
    if (u > range / 2) {
        u = range - u;
        S3 = -S3;
    }
    float p = __fmaf_rn((float)u, S1, S2);
    return __fmul_rn(S3, erfcinvf(__fadd_rn(p, p)));
}
__device__ __forceinline__ float minstd_uniform(uint32_t &x)
{
    return __fmul_rn((float)(minstd_next(x) - 1u), 4.656612873077392578125e-10f); // /(1+float(range)) == /2^31
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) with a (seed lo, seed hi) key: the device-side default start
// positions (init_positions_kernel below).  The FAST walk's own generators live in walk_fast.cuh.
// ------------------------------------------------------------------------------------------------
enum : uint32_t { STREAM_WALK = 0u, STREAM_PERMEABILITY = 1u, STREAM_XYZ0 = 2u };

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += W0;
        k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
}
// 23 random mantissa bits -> (0,1]  and  [0,1)
__device__ __forceinline__ float u01_open0(uint32_t r) { return 2.0f - __uint_as_float((r >> 9) | 0x3f800000u); }
__device__ __forceinline__ float u01_open1(uint32_t r) { return __uint_as_float((r >> 9) | 0x3f800000u) - 1.0f; }

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Ensemble sums in integer fixed point: a component is rounded to a multiple of 2^-22 once, every addition after that is exact, so the
// sums are order independent and bit-reproducible run to run (and across shards).  |M| < 16 is assumed (|M| <= 1 for M0 = (0,0,1)).
// May be called from divergent code: the lanes of a warp that arrive together with the same key (echo, substrate) are found with
// match.any and reduced with the warp's integer reduction unit; one lane per group adds to the block's shared-memory accumulators.
constexpr float kSumScale = 4194304.f; // 2^22
__device__ __forceinline__ void echo_sums_add(long long *b /* [entries][4] */, uint32_t key, const float *m)
{
    const unsigned peers = __match_any_sync(__activemask(), key);
    const int sx = __reduce_add_sync(peers, __float2int_rn(m[0] * kSumScale));
    const int sy = __reduce_add_sync(peers, __float2int_rn(m[1] * kSumScale));
    const int sz = __reduce_add_sync(peers, __float2int_rn(m[2] * kSumScale));
    if ((threadIdx.x & 31u) == (uint32_t)(__ffs(peers) - 1)) {
        unsigned long long *d = reinterpret_cast<unsigned long long *>(b) + (size_t)key * 4u;
        atomicAdd(d + 0, (unsigned long long)(long long)sx);
        atomicAdd(d + 1, (unsigned long long)(long long)sy);
        atomicAdd(d + 2, (unsigned long long)(long long)sz);
        atomicAdd(d + 3, (unsigned long long)__popc(peers));
    }
}

template <class T>
__device__ __forceinline__ const T *blob_ptr(const uint8_t *base, uint32_t off) { return reinterpret_cast<const T *>(base + off); }

// ------------------------------------------------------------------------------------------------
// The COMPAT kernel.  One thread = one (spin, scale).  grid = ceil(n_local/256) * n_scales blocks; consecutive
// blocks take consecutive SCALES of the same spin chunk so that memory-bound (small FoV scale) and
// issue-bound (large FoV scale) blocks share an SM.
// ------------------------------------------------------------------------------------------------
template <bool STATS>
__global__ void __launch_bounds__(kBlock) walk_compat_kernel(const WalkArgs A)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const BlobLayout &L = A.L;

    // ---- stage the sequence tables in shared memory -------------------------------------------
    const uint8_t *B = A.blob;
    uint32_t smem_used = 0;
    if (A.blob_in_smem) {
        const uint32_t nw = L.bytes / 4;
        const uint32_t *src = reinterpret_cast<const uint32_t *>(A.blob);
        uint32_t *dst = reinterpret_cast<uint32_t *>(smem);
        for (uint32_t i = threadIdx.x; i < nw; i += kBlock) dst[i] = __ldg(src + i);
        B = smem;
        smem_used = L.bytes;
    }
    // block-level sums [E][n_sub][4] (int64 fixed point) after the tables
    long long *bsum = reinterpret_cast<long long *>(smem + smem_used);
    const uint32_t n_bsum = A.sums_fx ? A.n_te * L.n_sub * 4u : 0u;
    for (uint32_t i = threadIdx.x; i < n_bsum; i += kBlock) bsum[i] = 0;
    __syncthreads();

    const int32_t  *tl_time = blob_ptr<int32_t>(B, L.tl_time);
    const uint32_t *tl_mask = blob_ptr<uint32_t>(B, L.tl_mask);
    const float *rf_s = blob_ptr<float>(B, L.rf_s), *rf_c = blob_ptr<float>(B, L.rf_c), *rf_ph = blob_ptr<float>(B, L.rf_ph);
    const float *deph_deg = blob_ptr<float>(B, L.deph_deg);
    const float *tgx = blob_ptr<float>(B, L.gx), *tgy = blob_ptr<float>(B, L.gy), *tgz = blob_ptr<float>(B, L.gz);
    const double *tsigma = blob_ptr<double>(B, L.sigma);
    const float *tT1 = blob_ptr<float>(B, L.T1s), *tT2 = blob_ptr<float>(B, L.T2s), *tpXY = blob_ptr<float>(B, L.pXY);

    // ---- which (spin, scale) ------------------------------------------------------------------
    const uint32_t k = blockIdx.x % A.n_scales;
    const uint32_t j = A.j_first + (blockIdx.x / A.n_scales) * kBlock + threadIdx.x; // thread slot in the shard
    bool alive = j < A.j_end;
    const uint32_t jl = alive ? (A.order ? A.order[j] : j) : 0u;          // local spin index
    const uint32_t spin_no = A.spin_first + jl;                          // GLOBAL spin id
    const float scale = __ldg(A.scales + k);

    float gscale = 1.f, lin_pc = A.lin_pc;
    if (A.scale_type == SWK_SCALE_GRADIENT) gscale = scale;
    else if (A.scale_type == SWK_SCALE_PHASE_CYCLING) lin_pc = __fmul_rn(A.lin_pc, scale); // monte_carlo.cu:303

    // ---- per-spin state: FP64 metres (kernels.cu:93-99) -----------------------------------------
    float m[3] = {0.f, 0.f, 1.f};
    float x0[3] = {0.f, 0.f, 0.f};
    if (alive) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            x0[i] = __ldg(A.xyz0 + 3 * (size_t)jl + i);
            if (A.m0) m[i] = __ldg(A.m0 + 3 * (size_t)jl + i);
        }
    }
    const int64_t nyz = (int64_t)A.ny * A.nz;
    double px[3] = {0, 0, 0}, fov_d[3] = {1, 1, 1}, s2g[3] = {0, 0, 0}, sigma_d = 0.;
    uint32_t rng_r = 1, rng_u = 1;
    int64_t ind_cur = 0;
    uint32_t ts_old = 0;
    {
        const uint32_t n3[3] = {A.nx, A.ny, A.nz};
#pragma unroll
        for (int i = 0; i < 3; i++) {
            float xs = (A.scale_type == SWK_SCALE_FOV) ? __fmul_rn(x0[i], scale) : x0[i];     // monte_carlo.cu:278
            float fv = (A.scale_type == SWK_SCALE_FOV) ? __fmul_rn(scale, A.fov[i]) : A.fov[i]; // monte_carlo.cu:280
            px[i] = (double)xs;
            fov_d[i] = (double)fv;
            s2g[i] = (double)n3[i] / fov_d[i]; // kernels.cu:98
        }
        int64_t ix = (int64_t)__dmul_rn(px[0], s2g[0]), iy = (int64_t)__dmul_rn(px[1], s2g[1]), iz = (int64_t)__dmul_rn(px[2], s2g[2]);
        ind_cur = ix * nyz + iy * (int64_t)A.nz + iz; // kernels.cu:102
        if (ind_cur < 0 || ind_cur >= A.V) { ind_cur = 0; alive = false; }
        if (alive) {
            ts_old = __ldg(A.mask + ind_cur);
            sigma_d = tsigma[ts_old];
            rng_r = rng_u = minstd_init(A.seed + spin_no); // kernels.cu:77-88: identical streams
        }
    }

    float field = 0.f, T1 = 0.f, T2 = 0.f; // kernels.cu:91
    float xyz_f[3];                        // last committed position as the reference stores it (float metres)
#pragma unroll
    for (int i = 0; i < 3; i++) xyz_f[i] = (float)px[i];

    unsigned long long st_mask = 0, st_field = 0, st_rej = 0, st_steps = 0;
    uint32_t itr = 0;
    bool lost = false;

    const size_t out_row = (size_t)k * A.n_local + jl;
    // echo slots + final position (WalkArgs::stage): slot e of this walker is stage[e * 32]
    const uint32_t srow = A.stage_by_slot ? j : jl;
    uint4 *stage = (A.stage && j < A.j_end) ? A.stage + (((size_t)k * A.stage_chunks + (srow >> 5)) * A.stage_row) * 32u + (srow & 31u) : nullptr;
    float *X1 = (A.record && A.XYZ1) ? A.XYZ1 + out_row * A.trj * 3 : nullptr;
    if (X1 && alive) { // slot 0 starts as the (scaled) initial position (kernels.cu:96)
        X1[0] = xyz_f[0]; X1[1] = xyz_f[1]; X1[2] = xyz_f[2];
    }

    const float rf_phase0 = A.rf_ph0;
    const uint32_t n_tp = A.n_tp;
    const float field_k = A.field_k;
    const bool has_field = A.fieldmap != nullptr;

    uint32_t echoes_done = 0; // echo events that fired in the last scan
    for (uint32_t scan = 0; scan < A.n_scans; scan++) {
        const bool last_scan = (scan + 1 == A.n_scans);
        // ---- phase cycling + first RF (kernels.cu:110-120) ----
        {
            float ph = (float)((double)(rf_phase0 + (float)scan * lin_pc) + (double)(scan * (scan + 1u)) / 2.0 * (double)A.quad_pc);
            while (ph > 360.0) ph = (float)(ph - 360.0);
            while (ph < 0) ph = (float)(ph + 360.0);
            float r[3];
            xrot_withphase(A.s, A.c, ph, m, r);
            m[0] = r[0]; m[1] = r[1]; m[2] = r[2];
        }
        bool fresh = true; // ind_old = matrix_length+1 (kernels.cu:123): first accepted step re-gathers
        uint32_t t = 0, t_old = 0;
        uint32_t cur_rf = 1, cur_te = 0, cnt_deph = 0, cnt_grad = 0;
        float acc = 0.f;

        for (uint32_t ev = 0; ev <= L.n_tl; ev++) {
            // walk until the step of timepoint `t_stop-1` has been accepted
            const uint32_t ev_time = ev < L.n_tl ? (uint32_t)tl_time[ev] : n_tp;
            const uint32_t t_stop = ev_time < n_tp ? ev_time + 1u : n_tp;

            // =============================== inner loop ===============================
            while (alive && t < t_stop) {
                double nx_d[3];
                // kernels.cu:130-137
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    nx_d[i] = px[i];
                    if (sigma_d != 0.) {
                        double rnd = __dmul_rn((double)minstd_normal(rng_r), sigma_d);
                        double xn = __dadd_rn(px[i], rnd);
                        if (xn < 0)
                            xn = __dadd_rn(xn, A.cross_fov ? fov_d[i] : __dadd_rn(fabs(rnd), fabs(rnd)));
                        else if (xn >= fov_d[i])
                            xn = __dadd_rn(xn, -(A.cross_fov ? fov_d[i] : __dadd_rn(fabs(rnd), fabs(rnd))));
                        nx_d[i] = xn;
                    }
                }
                const int64_t ix = (int64_t)__dmul_rn(nx_d[0], s2g[0]), iy = (int64_t)__dmul_rn(nx_d[1], s2g[1]),
                              iz = (int64_t)__dmul_rn(nx_d[2], s2g[2]);
                const int64_t ind_new = ix * nyz + iy * (int64_t)A.nz + iz; // kernels.cu:140
                if (ind_new >= A.V || ind_new < 0) {                        // kernels.cu:141-147
                    alive = false; lost = true;
                    break;
                }
                if (fresh || ind_new != ind_cur) { // kernels.cu:150-170
                    if (STATS) st_mask++;
                    uint32_t ts;
                    float fv;
                    if (A.raw_slab) { // the same two values from the [nx][ny] plane of a z-invariant phantom: one 8-byte gather (engine.cu raw_slab_kernel)
                        const uint2 w = __ldg(A.raw_slab + (ix * (int64_t)A.ny + iy));
                        ts = w.x;
                        fv = __uint_as_float(w.y);
                    } else {
                        ts = __ldg(A.mask + ind_new);
                        fv = has_field ? __ldg(A.fieldmap + ind_new) : 0.f; // issued together with the mask gather
                    }
                    if (ts != ts_old) {
                        if (minstd_uniform(rng_u) >= tpXY[ts_old * L.n_sub + ts]) {
                            if (STATS) st_rej++;
                            if (itr++ > A.max_iter) { alive = false; lost = true; break; }
                            continue; // redo the step from the old position; time does not advance
                        }
                        ts_old = ts;
                    }
                    if (STATS) st_field++;
                    ind_cur = ind_new;
                    fresh = false;
                    field = __fmul_rn(fv, field_k);       // monte_carlo.cu:244
                    T1 = tT1[ts_old];                     // kernels.cu:167-168 (ms -> s done on the host)
                    T2 = tT2[ts_old];
                    sigma_d = tsigma[ts_old];
                }
                acc += field; // kernels.cu:171-172
                itr = 0;
#pragma unroll
                for (int i = 0; i < 3; i++) px[i] = nx_d[i];
                if (A.record) { // kernels.cu:218-221
#pragma unroll
                    for (int i = 0; i < 3; i++) xyz_f[i] = (float)px[i];
                    if (X1) {
                        float *slot = X1 + 3 * ((size_t)scan * n_tp + t);
                        slot[0] = xyz_f[0]; slot[1] = xyz_f[1]; slot[2] = xyz_f[2];
                    }
                }
                if (STATS) st_steps++;
                t++;
            }
            // ============================ end of inner loop ============================
            if (ev >= L.n_tl || ev_time >= n_tp) { if (ev_time >= n_tp) break; else continue; }

            // ---- events of timepoint ev_time, in the reference's order (kernels.cu:175-215) ----
            const uint32_t mask_ev = tl_mask[ev];
            const uint32_t tp = ev_time;
            if (mask_ev & EV_DEPH) { // kernels.cu:175-178
                if (alive) acc += (float)spin_no * deph_deg[cnt_deph] / (float)A.n_spins_global;
                cnt_deph++;
            }
            if (mask_ev & EV_GRAD) { // kernels.cu:181-187
                if (alive) {
                    const float Gx = __fmul_rn(tgx[cnt_grad], gscale), Gy = __fmul_rn(tgy[cnt_grad], gscale),
                                Gz = __fmul_rn(tgz[cnt_grad], gscale); // monte_carlo.cu:288-290
                    // same association and contraction as the reference's SASS
                    double g = __fma_rn((double)Gz, px[2], __fma_rn((double)Gx, px[0], __dmul_rn((double)Gy, px[1])));
                    g = __dmul_rn(g, 1e-3);
                    g = __dmul_rn(g, (double)A.timestep_us);
                    g = __dmul_rn(g, 1e-6);
                    g = __dmul_rn(g, kGamma);
                    acc = (float)__fma_rn(g, kRad2Deg, (double)acc);
                }
                cnt_grad++;
            }
            if (mask_ev & EV_RF) { // kernels.cu:190-199
                if (alive) {
                    const float dt_s = (float)((double)((tp - t_old) * (uint32_t)A.timestep_us) * 1e-6);
                    dephase_relax(m, acc, T1, T2, dt_s);
                    float r[3];
                    xrot_withphase(rf_s[cur_rf], rf_c[cur_rf], rf_ph[cur_rf], m, r);
                    m[0] = r[0]; m[1] = r[1]; m[2] = r[2];
                    acc = 0.f;
                    t_old = tp;
                }
                cur_rf++;
            }
            if ((mask_ev & EV_ECHO) && last_scan) { // kernels.cu:202-215
                if (alive) {
                    const float dt_s = (float)((double)((tp - t_old) * (uint32_t)A.timestep_us) * 1e-6);
                    dephase_relax(m, acc, T1, T2, dt_s);
                    if (stage) stage[cur_te * 32u] = make_uint4(__float_as_uint(m[0]), __float_as_uint(m[1]), __float_as_uint(m[2]), ts_old);
                    acc = 0.f;
                    t_old = tp;
                    if (A.sums_fx) echo_sums_add(bsum, cur_te * L.n_sub + ts_old, m);
                } else if (stage) stage[cur_te * 32u] = make_uint4(0u, 0u, 0u, 0u); // abandoned spin: unwritten echoes read 0 (monte_carlo.cu:256,259-260)
                cur_te++;
            }
        }
        // ---- end of TR (kernels.cu:226-231) ----
        if (alive) {
            const float dt_s = (float)((double)((n_tp - t_old) * (uint32_t)A.timestep_us) * 1e-6);
            dephase_relax(m, acc, T1, T2, dt_s);
        }
        if (last_scan) echoes_done = cur_te;
    }

    // ---- final position (kernels.cu:220-221 leaves the last committed position in xyz1); echoes that never fired read 0 ----
    if (stage) {
        for (uint32_t e = echoes_done; e < A.n_te; e++) stage[e * 32u] = make_uint4(0u, 0u, 0u, 0u);
        if (!A.record) stage[A.n_te * 32u] = make_uint4(__float_as_uint((float)px[0]), __float_as_uint((float)px[1]), __float_as_uint((float)px[2]), lost ? 1u : 0u);
    }

    // ---- flush block sums and counters ----
    __syncthreads();
    if (A.sums_fx) {
        unsigned long long *gs = A.sums_fx + (size_t)k * n_bsum;
        for (uint32_t i = threadIdx.x; i < n_bsum; i += kBlock) {
            const long long v = bsum[i];
            if (v != 0) atomicAdd(gs + i, (unsigned long long)v);
        }
    }
    if (A.counters) {
        unsigned long long l = lost ? 1ull : 0ull;
        if (STATS) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                st_steps += __shfl_xor_sync(0xffffffffu, st_steps, o);
                st_mask += __shfl_xor_sync(0xffffffffu, st_mask, o);
                st_field += __shfl_xor_sync(0xffffffffu, st_field, o);
                st_rej += __shfl_xor_sync(0xffffffffu, st_rej, o);
            }
        }
        const unsigned lost_w = __popc(__ballot_sync(0xffffffffu, l != 0));
        if ((threadIdx.x & 31u) == 0) {
            if (STATS) {
                atomicAdd(A.counters + 0, st_steps);
                atomicAdd(A.counters + 1, st_mask);
                atomicAdd(A.counters + 2, st_field);
                atomicAdd(A.counters + 3, st_rej);
            }
            if (lost_w) atomicAdd(A.counters + 4, (unsigned long long)lost_w);
        }
    }
}

// Device-side default initial positions: uniform in [1%, 99%] of the FoV (distribution of
// monte_carlo.cu:142-151), Philox stream STREAM_XYZ0 keyed by (seed, global spin id).
__global__ void init_positions_kernel(float *xyz0, uint32_t n_local, uint32_t spin_first, uint64_t seed, float fx, float fy, float fz)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_local) return;
    const uint4 r = philox4x32_10(0u, 0u, spin_first + j, STREAM_XYZ0, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float f[3] = {fx, fy, fz};
    const uint32_t rr[3] = {r.x, r.y, r.z};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float a = (float)(0.01 * f[i]), b = (float)(0.99 * f[i]);
        xyz0[3 * (size_t)j + i] = u01_open1(rr[i]) * (b - a) + a;
    }
}

} // namespace swk
